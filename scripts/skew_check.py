"""Timing sanity on skewed key distributions (diagnostic)."""
import sys, torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 200_000_000
g = torch.Generator(device="cuda"); g.manual_seed(5)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
def run(name, k):
    ts = []
    for it in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        gb = engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)])
        s = gb.reduced(0); ng = gb.ngroups; gb.close()
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ok = abs(float(s.sum()) - float(v.sum())) < 1e-6 * float(v.sum())
    # one more call with the per-kernel-family profile on: where did the time go?
    engine.set_option("profile", 1); _lib.lib.dtb_profile_reset()
    gb = engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)]); gb.close()
    fam = {}
    for nm, ms in _lib.profile_records():
        fam[nm] = fam.get(nm, 0.0) + ms
    engine.set_option("profile", 0)
    split = " ".join(f"{a}={b:.2f}" for a, b in sorted(fam.items(), key=lambda t: -t[1])[:5])
    print(f"{name:28s} ngroups={ng:9d}  {min(ts):8.2f} ms  {n/min(ts)/1e6:8.2f} Grows/s  sums_ok={ok}  [{split}]", flush=True)
run("uniform 1e6 keys", torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32))
run("2 keys", torch.randint(0, 2, (n,), generator=g, device="cuda", dtype=torch.int32))
run("100 keys", torch.randint(0, 100, (n,), generator=g, device="cuda", dtype=torch.int32))
z = (torch.rand(n, generator=g, device="cuda") ** 8 * 1_000_000).to(torch.int32)      # heavy head
run("power-law head (x^8)", z)
hot = torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32); hot[::2] = 7
run("one key owns 50 %", hot)
sp = torch.randint(0, 300, (n,), generator=g, device="cuda", dtype=torch.int32) * 9973 + 11
run("300 keys in a 3e6 domain", sp)
run("sorted input", torch.sort(torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32)).values)
