"""Step time and per-kernel-family split on C2-shaped, skewed and float64-sort input (diagnostic)."""
import sys, torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
if len(sys.argv) > 2: engine.set_option("radix_bits", int(sys.argv[2]))
g = torch.Generator(device="cuda"); g.manual_seed(5)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)

def timed(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)

def families(fn):
    engine.set_option("profile", 1); _lib.lib.dtb_profile_reset(); fn()
    fam = {}
    for nm, ms in _lib.profile_records(): fam[nm] = fam.get(nm, 0.0) + ms
    engine.set_option("profile", 0)
    return " ".join(f"{a}={b:.2f}" for a, b in sorted(fam.items(), key=lambda t: -t[1])[:4])

def gb(k):
    def f():
        h = engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v[:k.numel()])]); h.close()
    return f
def so(k):
    def f():
        engine.Groupby([k], [4], 1).close()                    # SORT_ONLY: RowIndex only
    return f

cases = [("C2 uniform 1e6 keys", gb, torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32))]
m = n // 4
z = (torch.rand(m, generator=g, device="cuda") ** 8 * 1_000_000).to(torch.int32)
cases.append(("power-law (n/4)", gb, z))
cases.append(("2 keys (n/4)", gb, torch.randint(0, 2, (m,), generator=g, device="cuda", dtype=torch.int32)))
cases.append(("f64 sort (n/4)", so, torch.randn(m, generator=g, device="cuda", dtype=torch.float64)))
for name, mk, k in cases:
    f = mk(k)
    best, avg = timed(f)
    print(f"{name:22s} best {best:7.2f} ms  avg {avg:7.2f} ms  [{families(f)}]", flush=True)
