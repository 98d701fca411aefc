"""One C2-shaped groupby-sum call (for ncu captures).  Usage: one_step.py [rows] [groups]"""
import sys, torch
sys.path.insert(0, ".")
from datatable_b200 import engine, _lib
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 250_000_000
ng = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
g = torch.Generator(device="cuda"); g.manual_seed(5)
k = torch.randint(0, ng, (n,), generator=g, device="cuda", dtype=torch.int32)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
h = engine.Groupby([k], [0], 1, reducers=[(_lib.OP_SUM, v)])
print("ngroups", h.ngroups, "sum", float(h.reduced(0).sum()), float(v.sum()))
h.close()
