"""Host-side wall-clock vs device-side event time of one group()+reduce() call (diagnostic)."""
import sys, time
import torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(42)
k = torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
for it in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    order, offsets, ng = engine.group([k], [0], 1)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    s = engine.reduce(_lib.OP_SUM, v, order, offsets)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"iter {it}: group {1e3*(t1-t0):.2f} ms  reduce {1e3*(t2-t1):.2f} ms", flush=True)
engine.set_option("verbose", 2)
order, offsets, ng = engine.group([k], [0], 1)
