"""C2 / C3 step times with per-kernel-family profile records (CUDA events inside the engine)."""
import sys, torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
g = torch.Generator(device="cuda"); g.manual_seed(42)
k = torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)

def fam(fn):
    for _ in range(2): fn()
    engine.set_option("profile", 1); _lib.lib.dtb_profile_reset(); fn()
    out = {}
    for nm, ms in _lib.profile_records(): out.setdefault(nm, []).append(round(ms, 3))
    engine.set_option("profile", 0)
    return out

def timed(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

c2 = lambda: engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)]).close()
print("C2 %.3f ms" % timed(c2), fam(c2), flush=True)
gb = engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)])
s = gb.reduced(0); print("C2 check: ngroups", gb.ngroups, "sum err", abs(float(s.sum()) - float(v.sum())) / float(v.sum()))
off = torch.empty(gb.ngroups + 1, dtype=torch.int32, device="cuda")
engine._memcpy_d2d(off.data_ptr(), gb.offsets_ptr, off.numel() * 4)
cnt = torch.bincount(k, minlength=1_000_000)
print("C2 check: offsets == bincount prefix:", bool((off[1:].long() - off[:-1].long() == cnt[cnt > 0]).all()), flush=True)
gb.close(); del s, off, cnt
if len(sys.argv) > 2:
    del v
    x = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
    c3 = lambda: engine.Groupby([x], [4], 1).close()
    print("C3 %.3f ms" % timed(c3, 3), fam(c3), flush=True)
