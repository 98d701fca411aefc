import sys, torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib
n = 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(42)
k = torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
c2 = lambda: engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)])
for _ in range(3): c2().close()
torch.cuda.synchronize()
engine.set_option("verbose", 2)
import time
for _ in range(2):
    t0 = time.perf_counter(); gb = c2(); t1 = time.perf_counter(); s = gb.reduced(0); gb.close(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"python: Groupby() returned after {1e3*(t1-t0):.3f} ms, reduced+close+sync {1e3*(t2-t1):.3f} ms", flush=True)
