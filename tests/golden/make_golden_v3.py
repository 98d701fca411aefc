#!/usr/bin/env python
"""
Generates tests/golden/golden_v3.{npz,json} by running the *reference itself* (the unmodified build staged by
oracle/build_ref.sh) on the `i` node of DT[i, j, by(), sort()] when `i` is an integer or an integer slice:

    PYTHONPATH=oracle/_ref python tests/golden/make_golden_v3.py

    FExpr_Literal_SliceInt::evaluate_iby   (expr/fexpr_literal_sliceint.cc:82-170)  -- slice applied inside every group
    FExpr_Literal_Int::evaluate_iby        (expr/fexpr_literal_int.cc:146-192)      -- i-th row of every group
    EvalContext::evaluate                  (expr/eval_context.cc:154-163)           -- composed with the RowIndex of group()

Every case stores the key column, the slice, and what the reference returns for DT[i, :, by(k)] / DT[i, :, sort(k)] /
DT[i, {reducers}, by(k)]: the original row numbers in output order (column r = 0..n-1), the by column, and the sums /
counts per remaining group.  The reference cannot travel to the GPU box, so the vectors are committed.
"""
import json
import os

import numpy as np

import datatable as dt
from datatable import f, by, sort

HERE = os.path.dirname(os.path.abspath(__file__))
arrays, manifest = {}, []
rng = np.random.default_rng(20260924)
NA32 = -2**31


def frame(k, v):
    return dt.Frame(k=[None if x == NA32 else int(x) for x in k.tolist()], r=list(range(len(k))), v=v.tolist(),
                    stypes={"k": dt.int32, "r": dt.int32, "v": dt.float64})


def npcol(fr, name, dtype, na):
    lst = fr[:, name].to_list()[0]
    return np.array([na if x is None else x for x in lst], dtype=dtype)


def add(name, k, v, i, mode):
    DT = frame(k, v)
    sl = i if isinstance(i, int) else slice(*i)
    if mode == "by":
        R = DT[sl, :, by(f.k)]
    elif mode == "sort":
        R = DT[sl, :, sort(f.k)]
    else:
        R = DT[sl, {"s": dt.sum(f.v), "n": dt.count(f.v), "first_r": dt.first(f.r)}, by(f.k)]
    case = {"name": name, "mode": mode, "i": i if isinstance(i, int) else [None if x is None else int(x) for x in i],
            "nrows": int(R.nrows), "names": list(R.names)}
    arrays[name + ".k"] = k; arrays[name + ".v"] = v
    arrays[name + ".out_k"] = npcol(R, "k", np.int32, NA32)
    if mode == "red":
        arrays[name + ".out_s"] = npcol(R, "s", np.float64, np.nan)
        arrays[name + ".out_n"] = npcol(R, "n", np.int64, -2**63)
        arrays[name + ".out_first_r"] = npcol(R, "first_r", np.int32, NA32)
    else:
        arrays[name + ".out_r"] = npcol(R, "r", np.int32, NA32)
    manifest.append(case)


def keys(n, ng, na=0.1):
    k = rng.integers(0, ng, n).astype(np.int32)
    k[rng.random(n) < na] = NA32
    return k


n = 257
k1, v1 = keys(n, 9), np.round(rng.standard_normal(n), 3)
k2, v2 = np.array([5, 5, 5, 1, 1, 7, NA32, 7, 7, 7, 7, 3], dtype=np.int32), np.arange(12, dtype=np.float64)
slices = [(None, 2, None), (1, None, None), (None, None, 2), (2, 9, 3), (-3, None, None), (None, -2, None), (-4, -1, 2),
          (None, None, -1), (None, None, -2), (5, None, -1), (-2, None, -1), (5, 1, -2), (None, 2, -1), (-1, -5, -1),
          (100, None, None), (None, 0, None), (1, 3, 0), (-1, 2, 0), (-40, None, None), (None, None, -100), (3, 3, None)]
for j, s in enumerate(slices):
    add(f"iby.s{j}.rand", k1, v1, s, "by")
    add(f"iby.s{j}.lit", k2, v2, s, "by")
for j, s in enumerate(slices[:12]):
    add(f"isort.s{j}", k1, v1, s, "sort")
    add(f"ired.s{j}", k1, v1, s, "red")
for iv in (0, 1, 3, 30, -1, -2, -4, -50):
    add(f"iby.int{iv}.rand", k1, v1, iv, "by")
    add(f"iby.int{iv}.lit", k2, v2, iv, "by")
    add(f"ired.int{iv}", k1, v1, iv, "red")
# one group / every row its own group / empty frame
add("iby.onegroup", np.zeros(40, np.int32), np.arange(40.0), (3, None, 4), "by")
add("iby.allgroups", np.arange(40, dtype=np.int32)[::-1].copy(), np.arange(40.0), (None, 1, None), "by")
add("iby.empty", np.zeros(0, np.int32), np.zeros(0), (1, 5, 2), "by")

np.savez_compressed(os.path.join(HERE, "golden_v3.npz"), **arrays)
json.dump({"generator": "tests/golden/make_golden_v3.py", "datatable_version": dt.__version__, "cases": manifest},
          open(os.path.join(HERE, "golden_v3.json"), "w"), indent=0)
print(len(manifest), "cases")
