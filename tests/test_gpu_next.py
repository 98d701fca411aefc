"""GPU: the SURVEY.md 8(f) rows -- ordered reducers (first/last/sd/median/nunique), set operations,
mode/nmodal/nunique, keyed join -- through the C-ABI and the Frame mirror, against vectors produced by
the reference (tests/golden/golden_v2.*) and, at 1e6 rows, against the CPU oracle."""
import json
import os

import numpy as np
import pytest

from helpers import assert_reducer_equal, BOOL, INT8, INT16, INT32, INT64, FLOAT32, FLOAT64

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "tests", "golden")
META = json.load(open(os.path.join(G, "golden_v2.json")))
ARR = np.load(os.path.join(G, "golden_v2.npz"))
CASES = META["cases"]


def A(case, key):
    return ARR[f"{case['name']}__{key}"]


def by_kind(kind):
    return [c for c in CASES if c["kind"] == kind]


def eq(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and np.array_equal(a, b, equal_nan=a.dtype.kind == "f")


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("case", by_kind("ordered"), ids=[c["name"] for c in by_kind("ordered")])
def test_ordered_reducers_golden(case, device):
    import torch
    from datatable_b200 import engine, _lib
    OPS = {"first": _lib.OP_FIRST, "last": _lib.OP_LAST, "sd": _lib.OP_SD, "median": _lib.OP_MEDIAN, "nunique": _lib.OP_NUNIQUE}
    k = A(case, "k")
    kc = engine.Col(torch.from_numpy(k).cuda() if device else k, case["kst"])
    order, offsets, ng = engine.group([kc], [0], _lib.NA_FIRST)
    for i, (op, vi) in enumerate(case["reducers"]):
        v, vst = A(case, f"v{vi}"), case["vst"][vi]
        vc = engine.Col(torch.from_numpy(v).cuda() if device else v, vst)
        o = engine.sort_grouped(vc, order, offsets) if op in ("median", "nunique") else order
        got = engine.reduce(OPS[op], vc, o, offsets)
        got = got.cpu().numpy() if device else got
        want = A(case, f"red{i}")
        if op in ("first", "last"):
            assert eq(got, want), f"{case['name']} {op}(v{vi})"
        else:
            assert_reducer_equal(got, want, "mean" if op in ("sd", "median") else "count", vst, f"{case['name']} {op}(v{vi})")


def test_ordered_reducers_frame_api():
    import datatable_b200 as dtb
    from datatable_b200 import f, by
    case = by_kind("ordered")[6]
    k = A(case, "k")
    cols = {"k": k}
    for vi in range(7):
        cols[f"v{vi}"] = A(case, f"v{vi}")
    DT = dtb.Frame(cols, stypes={"k": case["kst"], **{f"v{vi}": case["vst"][vi] for vi in range(7)}})
    fn = {"first": dtb.first, "last": dtb.last, "sd": dtb.sd, "median": dtb.median, "nunique": dtb.nunique}
    for i, (op, vi) in enumerate(case["reducers"]):
        R = DT[:, {"r": fn[op](f[f"v{vi}"])}, by(f.k)]
        got, want = R.to_numpy("r"), A(case, f"red{i}")
        if op in ("first", "last"):
            assert eq(got, want)
        else:
            assert_reducer_equal(got, want, "mean" if op in ("sd", "median") else "count", case["vst"][vi], f"{op}(v{vi})")
    # first/last keep the column's stype (bool8 stays bool8); sd/median of ints are float64
    R = DT[:, {"a": dtb.first(f.v0), "b": dtb.sd(f.v1), "c": dtb.nunique(f.v6), "m": dtb.median(f.v5)}, by(f.k)]
    assert R.stypes[1:] == (BOOL, FLOAT64, INT64, FLOAT32)


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("case", by_kind("sets"), ids=[c["name"] for c in by_kind("sets")])
def test_set_operations_golden(case, device):
    import datatable_b200 as dtb
    st, K = case["st"], case["K"]
    frames = []
    for i in range(K):
        fr = dtb.Frame({"A": A(case, f"in{i}")}, stypes={"A": st})
        frames.append(fr.to_device() if device else fr)
    fns = {"union": dtb.union, "intersect": dtb.intersect, "setdiff": dtb.setdiff, "symdiff": dtb.symdiff}
    for op in case["ops"]:
        R = fns[op](*frames)
        assert R.names == ("A",) and R.stypes == (st,)
        assert eq(R.to_numpy("A"), A(case, op)), f"{case['name']} {op}"
    if case.get("no_stats"):
        return
    F0 = frames[0]
    assert eq(dtb.unique(F0).to_numpy("A"), A(case, "unique0"))
    assert eq(F0.nunique().to_numpy("A"), A(case, "nunique0"))
    assert eq(F0.nmodal().to_numpy("A"), A(case, "nmodal0"))
    assert eq(F0.mode().to_numpy("A"), A(case, "mode0"))


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("case", by_kind("join"), ids=[c["name"] for c in by_kind("join")])
def test_keyed_join_golden(case, device):
    import datatable_b200 as dtb
    from datatable_b200 import join
    nk = len(case["xst"])
    names = [f"k{i}" for i in range(nk)]
    J = dtb.Frame({nm: A(case, f"jraw{i}") for i, nm in enumerate(names)}, stypes=dict(zip(names, case["jst"])))
    X = dtb.Frame({nm: A(case, f"x{i}") for i, nm in enumerate(names)}, stypes=dict(zip(names, case["xst"])))
    if device:
        J, X = J.to_device(), X.to_device()
    J.key = names                                                     # sorts J, checks uniqueness
    for i, nm in enumerate(names):
        assert eq(J.to_numpy(nm), A(case, f"jsorted{i}")), "key columns after J.key = ..."
    J2 = dtb.Frame({**{nm: J.column(nm) for nm in names}, "jrow": np.arange(J.nrows, dtype=np.int32)},
                   stypes=dict(zip(names, case["jst"])))
    if device:
        J2 = J2.to_device()
    J2.key = names
    R = X[:, :, join(J2)]
    assert R.names == tuple(names) + ("jrow",) and R.nrows == X.nrows
    assert eq(R.to_numpy("jrow"), A(case, "index")), case["name"]


def test_key_must_be_unique():
    import datatable_b200 as dtb
    F = dtb.Frame({"k": np.array([1, 2, 2], np.int32)})
    with pytest.raises(ValueError, match="Cannot set a key: the values are not unique"):
        F.key = "k"
    with pytest.raises(ValueError, match="not keyed"):
        dtb.join(F)


def test_large_vs_oracle():
    """1e6 rows: ordered reducers, set selection, largest group and join against the CPU oracle."""
    import torch
    from datatable_b200 import engine, _lib
    from oracle import oracle as orc
    rng = np.random.default_rng(77)
    n = 1_000_000
    k = rng.integers(0, 5000, n).astype(np.int32)
    v = rng.standard_normal(n); v[rng.random(n) < 0.05] = np.nan
    w = rng.integers(-50, 50, n).astype(np.int16); w[rng.random(n) < 0.05] = -2**15
    wo, wf, wng = orc.group([k], [0], orc.NA_FIRST)
    kd, vd, wd = (torch.from_numpy(a).cuda() for a in (k, v, w))
    gb = engine.Groupby([kd], [0], _lib.NA_FIRST)
    assert np.array_equal(gb.order().cpu().numpy(), wo)
    for col, cold, st in ((v, vd, FLOAT64), (w, wd, INT16)):
        for name, op, oop in (("first", _lib.OP_FIRST, orc.FIRST), ("last", _lib.OP_LAST, orc.LAST), ("sd", _lib.OP_SD, orc.SD)):
            got = gb.reduce(op, cold).cpu().numpy()
            want = orc.reduce(oop, col, wo, wf)
            if name == "sd":
                assert_reducer_equal(got, want, "mean", st, name)
            else:
                assert eq(got, want), name
        o2 = gb.sort_grouped(cold)
        assert np.array_equal(o2.cpu().numpy(), orc.sort_grouped(col, wo, wf))
        for name, op, oop in (("median", _lib.OP_MEDIAN, orc.MEDIAN), ("nunique", _lib.OP_NUNIQUE, orc.NUNIQUE)):
            got = gb.reduce_ordered(op, cold, o2).cpu().numpy()
            want = orc.reduce(oop, col, o2.cpu().numpy(), wf)
            assert_reducer_equal(got, want, "mean" if name == "median" else "count", st, name)
    idx, size = engine.largest_group(gb.offsets(), 0)
    assert (idx, size) == orc.largest_group(wf, 0)
    gb.close()
    # set selection over three concatenated inputs
    ins = [rng.integers(0, 3000, m).astype(np.int32) for m in (200_000, 150_000, 90_000)]
    cat = np.concatenate(ins); cs = np.cumsum([len(a) for a in ins])
    o, f_, ng = orc.group([cat], [0], orc.NA_FIRST)
    od, fd, _ = engine.group([torch.from_numpy(cat).cuda()], [0], _lib.NA_FIRST)
    for mode in (orc.SET_UNION, orc.SET_INTERSECT, orc.SET_SETDIFF, orc.SET_SYMDIFF):
        assert np.array_equal(engine.set_select(mode, od, fd, cs).cpu().numpy(), orc.set_select(mode, o, f_, cs))
    # join: 2e5 X rows against 5e4 unique sorted keys
    jk = np.unique(rng.integers(-10**6, 10**6, 50_000)).astype(np.int64)
    xk = rng.integers(-10**6, 10**6, 200_000).astype(np.int32)
    got = engine.join_index([torch.from_numpy(xk).cuda()], [torch.from_numpy(jk).cuda()]).cpu().numpy()
    pos = np.searchsorted(jk, xk.astype(np.int64))
    pos_c = np.minimum(pos, len(jk) - 1)
    want = np.where(jk[pos_c] == xk, pos_c, -2**31).astype(np.int32)
    assert np.array_equal(got, want)
