"""One profiled call of a named case between cudaProfilerStart/Stop (run under `ncu --profile-from-start off`)."""
import sys
import torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib

case = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 250_000_000
for kv in sys.argv[3:]:
    name, val = kv.split("=")
    engine.set_option(name, int(val))
g = torch.Generator(device="cuda"); g.manual_seed(5)
if case == "c2":
    k = torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32)
    v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
    def fn():
        engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)]).close()
elif case == "c3":
    x = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    def fn():
        engine.Groupby([x], [4], 1).close()
elif case == "c4":
    k1 = torch.randint(0, 1000, (n,), generator=g, device="cuda", dtype=torch.int64) << 33
    k2 = torch.randint(0, 1000, (n,), generator=g, device="cuda", dtype=torch.int32)
    vs = [torch.randn(n, generator=g, device="cuda", dtype=torch.float64) for _ in range(3)]
    ops = [_lib.OP_MEAN, _lib.OP_MIN, _lib.OP_MAX, _lib.OP_COUNT]
    def fn():
        engine.Groupby([k1, k2], [0, 0], 1, reducers=[(op, v) for v in vs for op in ops]).close()
fn(); fn()
torch.cuda.synchronize()
torch.cuda.profiler.start()
fn()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", case, n)
