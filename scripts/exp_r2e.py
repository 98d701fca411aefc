"""Experiment: bucketed multi-reducer variants on the C4 shape (and an int32 value column), profile records
per kernel family, results compared with the one-atomic-per-row path."""
import sys, torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
k1 = torch.randint(0, 1000, (n,), generator=g, device="cuda", dtype=torch.int64) << 33
k2 = torch.randint(0, 1000, (n,), generator=g, device="cuda", dtype=torch.int32)
vs = [torch.randn(n, generator=g, device="cuda", dtype=torch.float64) for _ in range(3)]
for v in vs: v[::100] = float("nan")
vi = torch.randint(-1000, 1000, (n,), generator=g, device="cuda", dtype=torch.int32)
vi[::77] = -2**31
ops = [_lib.OP_MEAN, _lib.OP_MIN, _lib.OP_MAX, _lib.OP_COUNT]
reds = [(op, v) for v in vs for op in ops] + [(op, vi) for op in (_lib.OP_SUM, _lib.OP_MEAN, _lib.OP_MIN, _lib.OP_MAX, _lib.OP_COUNT, _lib.OP_COUNTNA)]

def run(keep=False, nred=len(reds)):
    gb = engine.Groupby([k1, k2], [0, 0], _lib.NA_FIRST, reducers=reds[:nred])
    res = [gb.reduced(i) for i in range(nred)] if keep else None
    gb.close()
    return res

def fam(fn):
    for _ in range(2): fn()
    engine.set_option("profile", 1); _lib.lib.dtb_profile_reset(); fn()
    out = {}
    for nm, ms in _lib.profile_records(): out.setdefault(nm, []).append(round(ms, 3))
    engine.set_option("profile", 0)
    return out

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

engine.set_option("bucketed_reducers", 0)
ref = run(keep=True)
print("direct path: C4 (12 reducers) %.2f ms" % timed(lambda: run(nred=12)), flush=True)
engine.set_option("bucketed_reducers", 1)
for var in (0,):
    res = run(keep=True)
    bad = []
    for i, (a, b) in enumerate(zip(ref, res)):
        if a.dtype.is_floating_point:
            same = torch.allclose(a, b, rtol=1e-9, atol=1e-9, equal_nan=True) if reds[i][0] in (_lib.OP_MEAN, _lib.OP_SUM) else \
                   bool(((a == b) | (torch.isnan(a) & torch.isnan(b))).all())
        else:
            same = bool((a == b).all())
        if not same: bad.append(i)
    print("variant", var, "mismatching reducers:", bad, flush=True)
    print("variant", var, "C4 (12 reducers) %.2f ms ; with int32 column (18 reducers) %.2f ms" %
          (timed(lambda: run(nred=12)), timed(lambda: run())), flush=True)
    pr = fam(lambda: run(nred=12))
    print("   ", {k: v for k, v in pr.items() if "bucket" in k or "compose" in k}, flush=True)
