"""CPU: the oracle's restatements of the SURVEY.md 8(f) rows (ordered reducers, set operations, mode,
keyed join) against vectors produced by the reference itself (tests/golden/make_golden_v2.py)."""
import json
import os

import numpy as np
import pytest

from helpers import assert_reducer_equal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
META = json.load(open(os.path.join(G, "golden_v2.json")))
ARR = np.load(os.path.join(G, "golden_v2.npz"))
CASES = META["cases"]


def A(case, key):
    return ARR[f"{case['name']}__{key}"]


def by_kind(kind):
    return [c for c in CASES if c["kind"] == kind]


@pytest.mark.parametrize("case", by_kind("ordered"), ids=[c["name"] for c in by_kind("ordered")])
def test_ordered_reducers(case):
    from oracle import oracle as orc
    OPS = {"first": orc.FIRST, "last": orc.LAST, "sd": orc.SD, "median": orc.MEDIAN, "nunique": orc.NUNIQUE}
    k = A(case, "k")
    order, offsets, ng = orc.group([k], [0], orc.NA_FIRST, stypes=[case["kst"]])
    for i, (op, vi) in enumerate(case["reducers"]):
        v, vst = A(case, f"v{vi}"), case["vst"][vi]
        o = orc.sort_grouped(v, order, offsets, stype=vst) if op == "median" else order
        got = orc.reduce(OPS[op], v, o, offsets, stype=vst)
        want = A(case, f"red{i}")
        if op in ("first", "last"):
            assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=got.dtype.kind == "f"), f"{case['name']} {op}(v{vi})"
        else:
            assert_reducer_equal(got, want, "mean" if op in ("sd", "median") else "count", vst, f"{case['name']} {op}(v{vi})")


def _set_op(orc, mode, ins, st):
    cat = np.concatenate(ins)
    order, offsets, ng = orc.group([cat], [0], orc.NA_FIRST, stypes=[st])
    rows = orc.set_select(mode, order, offsets, np.cumsum([len(a) for a in ins]))
    return cat[rows]


@pytest.mark.parametrize("case", by_kind("sets"), ids=[c["name"] for c in by_kind("sets")])
def test_set_operations_and_mode(case):
    from oracle import oracle as orc
    st, K = case["st"], case["K"]
    ins = [A(case, f"in{i}") for i in range(K)]
    modes = {"union": orc.SET_UNION, "intersect": orc.SET_INTERSECT, "setdiff": orc.SET_SETDIFF, "symdiff": orc.SET_SYMDIFF}
    for op in case["ops"]:
        got = _set_op(orc, modes[op], ins, st) if sum(len(a) for a in ins) else np.zeros(0, ins[0].dtype)
        want = A(case, op)
        assert np.array_equal(got, want, equal_nan=got.dtype.kind == "f"), f"{case['name']} {op}"
    if case.get("no_stats"):
        return
    a = ins[0]
    order, offsets, ng = orc.group([a], [0], orc.NA_FIRST, stypes=[st])
    assert np.array_equal(a[order[offsets[:-1]]], A(case, "unique0"), equal_nan=a.dtype.kind == "f")
    has_na = bool(orc._na_mask(a[order[:1]], st)[0]) if len(a) else False
    assert ng - int(has_na) == int(A(case, "nunique0")[0])                    # stats.cc:977-979
    idx, size = orc.largest_group(offsets, int(has_na))
    assert size == int(A(case, "nmodal0")[0])
    if size:
        assert np.array_equal(a[order[offsets[idx]]:order[offsets[idx]] + 1], A(case, "mode0"), equal_nan=True)


@pytest.mark.parametrize("case", by_kind("join"), ids=[c["name"] for c in by_kind("join")])
def test_keyed_join(case):
    from oracle import oracle as orc
    nk = len(case["xst"])
    # setting the key = group() on the key columns, uniqueness check, reorder (frame/key.cc:118-180)
    jraw = [A(case, f"jraw{i}") for i in range(nk)]
    if len(jraw[0]):
        order, offsets, ng = orc.group(jraw, [0] * nk, orc.NA_FIRST, stypes=case["jst"])
        assert ng == len(jraw[0])
        jsorted = [a[order] for a in jraw]
    else:
        jsorted = jraw
    for i in range(nk):
        assert np.array_equal(jsorted[i], A(case, f"jsorted{i}"), equal_nan=True)
    got = orc.join_index([A(case, f"x{i}") for i in range(nk)], case["xst"], jsorted, case["jst"])
    assert np.array_equal(got, A(case, "index")), case["name"]


def test_duplicate_key_error_text():
    assert META["duplicate_key_error"] == "ValueError: Cannot set a key: the values are not unique"
