"""Round-2 experiment C (diagnostic): C4 with and without the bucketed multi-reducer; per-kernel split."""
import sys
import torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
k1 = torch.randint(0, 1000, (n,), generator=g, device="cuda", dtype=torch.int64) << 33
k2 = torch.randint(0, 1000, (n,), generator=g, device="cuda", dtype=torch.int32)
vs = [torch.randn(n, generator=g, device="cuda", dtype=torch.float64) for _ in range(3)]
for v in vs: v[::100] = float("nan")
ops = [_lib.OP_MEAN, _lib.OP_MIN, _lib.OP_MAX, _lib.OP_COUNT]
def c4(keep=False):
    gb = engine.Groupby([k1, k2], [0, 0], 1, reducers=[(op, v) for v in vs for op in ops])
    res = [gb.reduced(i) for i in range(12)] if keep else None
    gb.close(); return res
def timed(fn, reps=3, warm=2):
    for _ in range(warm): fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return min(ts)
def families(fn):
    engine.set_option("profile", 1); _lib.lib.dtb_profile_reset(); fn()
    fam = {}
    for nm, ms in _lib.profile_records(): fam.setdefault(nm, []).append(ms)
    engine.set_option("profile", 0)
    return " ".join(f"{a}={sum(b):.2f}({len(b)})" for a, b in sorted(fam.items(), key=lambda t: -sum(t[1])))
ref = None
for b in (0, 1):
    engine.set_option("bucketed_reducers", b)
    r = c4(keep=True)
    if ref is None: ref = r
    else:
        for i, (x, y) in enumerate(zip(ref, r)):
            ok = torch.allclose(x, y, rtol=1e-9, equal_nan=True) if x.dtype.is_floating_point else torch.equal(x, y)
            assert ok, f"reducer {i} differs between plain and bucketed"
    print(f"C4 n={n} bucketed={b}: {timed(c4):.2f} ms [{families(c4)}]", flush=True)
print("results identical")
