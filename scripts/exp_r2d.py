"""Diagnostic: why is the first scatter pass of a float64 sort slower than the later identical launches?"""
import os, sys, torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
g = torch.Generator(device="cuda"); g.manual_seed(5)
x = torch.randn(n, generator=g, device="cuda", dtype=torch.float64)
k = torch.randint(0, 1_000_000, (n,), generator=g, device="cuda", dtype=torch.int32)
def fam(fn):
    for _ in range(2): fn()
    engine.set_option("profile", 1); _lib.lib.dtb_profile_reset(); fn()
    out = {}
    for nm, ms in _lib.profile_records(): out.setdefault(nm, []).append(round(ms, 3))
    engine.set_option("profile", 0)
    return out
print("f64 sort", fam(lambda: engine.Groupby([x], [4], 1).close()))
print("i32 sort", fam(lambda: engine.Groupby([k], [4], 1).close()))
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
for sk in (1, 0):
    engine.set_option("stage_keys", sk)
    print("stage_keys", sk, "C2", fam(lambda: engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)]).close()))
    print("stage_keys", sk, "f64 sort", fam(lambda: engine.Groupby([x], [4], 1).close()))
engine.set_option("stage_keys", 1)
