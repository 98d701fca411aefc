#!/usr/bin/env python
"""
Generates tests/golden/golden_v2.{npz,json} by running the *reference itself* (the unmodified build
staged by oracle/build_ref.sh) on the SURVEY.md 8(f) rows:

    PYTHONPATH=oracle/_ref python tests/golden/make_golden_v2.py

    ordered reducers  dt.first / dt.last / dt.sd / dt.median / dt.nunique under by()
                      (expr/head_reduce_unary.cc:120-560)
    set operations    dt.unique / union / intersect / setdiff / symdiff   (set_funcs.cc:126-456)
    mode / nmodal     Frame.mode(), Frame.nmodal(), Frame.nunique()       (stats.cc:955-1003)
    keyed join        X[:, :, join(J)] with J.key set                     (frame/key.cc:118-180, frame/join.cc:392-470)

Each case stores its inputs and the reference's outputs.  The reference cannot travel to the GPU
box, so the vectors are committed.
"""
import json
import os

import numpy as np

import datatable as dt
from datatable import f, by, join

HERE = os.path.dirname(os.path.abspath(__file__))
BOOL, INT8, INT16, INT32, INT64, FLOAT32, FLOAT64 = 1, 2, 3, 4, 5, 6, 7
NP2ST = {np.dtype(np.int8): INT8, np.dtype(np.int16): INT16, np.dtype(np.int32): INT32,
         np.dtype(np.int64): INT64, np.dtype(np.float32): FLOAT32, np.dtype(np.float64): FLOAT64}
NA = {INT8: -2**7, INT16: -2**15, INT32: -2**31, INT64: -2**63, BOOL: -128}
DTST = {BOOL: dt.bool8, INT8: dt.int8, INT16: dt.int16, INT32: dt.int32, INT64: dt.int64,
        FLOAT32: dt.float32, FLOAT64: dt.float64}
arrays, manifest = {}, []
rng = np.random.default_rng(20260923)


def col(a, st, name):
    """single-column dt.Frame from a numpy array with NA sentinels"""
    if st in (FLOAT32, FLOAT64):
        fr = dt.Frame(np.ascontiguousarray(a))
    else:
        lst = [None if x == NA[st] else (bool(x) if st == BOOL else int(x)) for x in a.tolist()]
        fr = dt.Frame(lst, stype=DTST[st]) if len(lst) else dt.Frame(np.ascontiguousarray(a.astype(
            {BOOL: np.bool_, INT8: np.int8, INT16: np.int16, INT32: np.int32, INT64: np.int64}[st])))
    fr.names = [name]
    return fr


def to_np(fr, st=None):
    """reference column -> numpy with NA sentinels (bool -> int8)"""
    if fr.nrows == 0:
        return np.zeros(0, dtype={BOOL: np.int8, INT8: np.int8, INT16: np.int16, INT32: np.int32, INT64: np.int64,
                                  FLOAT32: np.float32, FLOAT64: np.float64}.get(st, np.float64))
    lst = fr.to_list()[0]
    stype = fr.stypes[0]
    if stype in (dt.float32, dt.float64):
        return np.array([np.nan if x is None else x for x in lst], dtype=np.float32 if stype == dt.float32 else np.float64)
    if stype == dt.bool8:
        return np.array([-128 if x is None else int(x) for x in lst], dtype=np.int8)
    npdt = {dt.int8: np.int8, dt.int16: np.int16, dt.int32: np.int32, dt.int64: np.int64}[stype]
    return np.array([np.iinfo(npdt).min if x is None else x for x in lst], dtype=npdt)


def rnd(st, n, spread="few", nafrac=0.1):
    if st == BOOL:
        a = rng.integers(0, 2, n).astype(np.int8)
    elif st in (FLOAT32, FLOAT64):
        dtp = np.float32 if st == FLOAT32 else np.float64
        a = (rng.integers(-6, 7, n) / 2).astype(dtp) if spread == "few" else rng.standard_normal(n).astype(dtp)
        if spread == "few" and n > 4:
            a[rng.integers(0, n, 2)] = -0.0
    else:
        dtp = {INT8: np.int8, INT16: np.int16, INT32: np.int32, INT64: np.int64}[st]
        lo, hi = (-5, 6) if spread == "few" else (max(np.iinfo(dtp).min + 1, -10**6), min(np.iinfo(dtp).max, 10**6))
        a = rng.integers(lo, hi, n).astype(dtp)
    m = rng.random(n) < nafrac
    if st in (FLOAT32, FLOAT64):
        a[m] = np.nan
    else:
        a[m] = NA[st]
    return a


# ---------------------------------------------------------------------------
# ordered reducers under by()
# ---------------------------------------------------------------------------
RED = {"first": dt.first, "last": dt.last, "sd": dt.sd, "median": dt.median, "nunique": dt.nunique}
cid = 0
for kst in (INT32, INT8, FLOAT64):
    for n in (1, 7, 300, 4000):
        k = rnd(kst, n, "few", 0.05)
        name = f"or{cid:02d}"; cid += 1
        meta = {"name": name, "kind": "ordered", "n": n, "kst": kst, "vst": [], "reducers": []}
        DT = col(k, kst, "k")
        arrays[f"{name}__k"] = k
        vals = []
        for vi, vst in enumerate((BOOL, INT8, INT16, INT32, INT64, FLOAT32, FLOAT64)):
            v = rnd(vst, n, "few" if vi % 2 == 0 else "wide", 0.25)
            arrays[f"{name}__v{vi}"] = v
            meta["vst"].append(vst)
            DT = dt.cbind(DT, col(v, vst, f"v{vi}"))
        for vi in range(7):
            for op in RED:
                R = DT[:, {"r": RED[op](f[f"v{vi}"])}, by(f.k)]
                arrays[f"{name}__red{len(meta['reducers'])}"] = to_np(R[:, "r"])
                meta["reducers"].append([op, vi])
        manifest.append(meta)

# ---------------------------------------------------------------------------
# set operations + unique/nunique/mode
# ---------------------------------------------------------------------------
SETS = {"union": dt.union, "intersect": dt.intersect, "setdiff": dt.setdiff, "symdiff": dt.symdiff}
for st in (INT32, INT64, FLOAT64, INT8, BOOL, FLOAT32):
    for K, n in ((1, 50), (2, 40), (2, 600), (3, 30), (4, 500)):
        name = f"st{cid:02d}"; cid += 1
        ins = [rnd(st, max(1, n + 13 * i), "few", 0.1) for i in range(K)]
        if K >= 3:
            ins[1] = ins[1][: max(1, len(ins[1]) // 3)]
        meta = {"name": name, "kind": "sets", "st": st, "K": K, "ops": list(SETS)}
        frames = []
        for i, a in enumerate(ins):
            arrays[f"{name}__in{i}"] = a
            frames.append(col(a, st, "A"))
        for op, fn in SETS.items():
            arrays[f"{name}__{op}"] = to_np(fn(*frames), st)
        arrays[f"{name}__unique0"] = to_np(dt.unique(frames[0]), st)
        F0 = frames[0]
        arrays[f"{name}__nunique0"] = np.array(F0.nunique().to_list()[0], dtype=np.int64)
        mode = F0.mode()
        arrays[f"{name}__mode0"] = to_np(mode, st)
        arrays[f"{name}__nmodal0"] = np.array(F0.nmodal().to_list()[0], dtype=np.int64)
        manifest.append(meta)
# empty inputs
name = f"st{cid:02d}"; cid += 1
e0 = np.zeros(0, np.int32); e1 = np.array([3, 1, 3], np.int32)
arrays[f"{name}__in0"] = e0; arrays[f"{name}__in1"] = e1
for op, fn in SETS.items():
    arrays[f"{name}__{op}"] = to_np(fn(col(e0, INT32, "A"), col(e1, INT32, "A")), INT32)
manifest.append({"name": name, "kind": "sets", "st": INT32, "K": 2, "ops": list(SETS), "no_stats": True})

# ---------------------------------------------------------------------------
# keyed join
# ---------------------------------------------------------------------------
def join_case(name, xkeys, xst, jkeys, jst, nj_payload=True):
    """J gets its key set (sorted, unique); X is joined; the golden is the J row matched by every X row,
    recovered from a payload column holding J's (sorted) row number."""
    J = None
    for i, (a, st) in enumerate(zip(jkeys, jst)):
        c = col(a, st, f"k{i}")
        J = c if J is None else dt.cbind(J, c)
    J.key = [f"k{i}" for i in range(len(jkeys))]
    J = dt.cbind(J, dt.Frame(jrow=np.arange(J.nrows, dtype=np.int32)))      # row number AFTER the key sort
    J.key = [f"k{i}" for i in range(len(jkeys))]
    X = None
    for i, (a, st) in enumerate(zip(xkeys, xst)):
        c = col(a, st, f"k{i}")
        X = c if X is None else dt.cbind(X, c)
    R = X[:, :, join(J)]
    for i, a in enumerate(xkeys):
        arrays[f"{name}__x{i}"] = a
    for i in range(len(jkeys)):
        arrays[f"{name}__jraw{i}"] = jkeys[i]
        arrays[f"{name}__jsorted{i}"] = to_np(J[:, f"k{i}"], jst[i])
    arrays[f"{name}__index"] = to_np(R[:, "jrow"], INT32)
    manifest.append({"name": name, "kind": "join", "xst": list(xst), "jst": list(jst)})


def uniq_rows(cols, sts):
    """drop duplicate key tuples so that J.key can be set"""
    fr = None
    for i, (a, st) in enumerate(zip(cols, sts)):
        c = col(a, st, f"k{i}")
        fr = c if fr is None else dt.cbind(fr, c)
    U = fr[:, dt.count(), by(*[f[f"k{i}"] for i in range(len(cols))])]
    return [to_np(U[:, f"k{i}"], st) for i, st in enumerate(sts)]


jc = 0
for xst, jst in (([INT32], [INT32]), ([INT64], [INT32]), ([INT8], [INT64]), ([FLOAT64], [INT32]), ([INT32], [FLOAT64]),
                 ([FLOAT64], [FLOAT64]), ([FLOAT32], [FLOAT64]), ([BOOL], [BOOL]), ([INT16], [INT8]),
                 ([INT32, INT32], [INT32, INT32]), ([INT64, FLOAT64], [INT32, FLOAT64]), ([INT8, INT16, INT32], [INT8, INT16, INT32])):
    for nx, nj in ((40, 12), (500, 200)):
        jk = uniq_rows([rnd(s, nj, "few", 0.15) for s in jst], jst)
        perm = rng.permutation(len(jk[0]))
        jk = [a[perm] for a in jk]                                      # J arrives unsorted; setting the key sorts it
        xk = [rnd(s, nx, "few", 0.15) for s in xst]
        for c in range(len(xst)):
            if xst[c] in (FLOAT32, FLOAT64) and jst[c] not in (FLOAT32, FLOAT64):
                xk[c][::5] = 0.25                                       # fractions never match an integer key
            if xst[c] == INT64 and jst[c] == INT32:
                xk[c][::7] = 2**40                                      # out of J's range
        join_case(f"jn{jc:02d}", xk, xst, jk, jst); jc += 1
# empty J, and a failing key (duplicate values)
join_case(f"jn{jc:02d}", [np.array([1, 2, -2**31], np.int32)], [INT32], [np.zeros(0, np.int32)], [INT32]); jc += 1
try:
    F = dt.Frame(k=[1, 2, 2]); F.key = "k"
    keyerr = None
except Exception as e:                                                  # noqa: BLE001
    keyerr = f"{type(e).__name__}: {e}"

np.savez_compressed(os.path.join(HERE, "golden_v2.npz"), **arrays)
with open(os.path.join(HERE, "golden_v2.json"), "w") as fh:
    json.dump({"reference": "h2oai/datatable @ 3611640 (1.2.0a), unmodified build (oracle/build_ref.sh)",
               "datatable_version": dt.__version__, "duplicate_key_error": keyerr, "cases": manifest}, fh, indent=0)
print(f"{len(manifest)} cases, {sum(a.nbytes for a in arrays.values())/1e6:.2f} MB raw; key error: {keyerr}")
