"""GPU: the `i` node of DT[i, j, by(), sort()] for integers and integer slices (dtb_slice_groups) against the
oracle's restatement and against vectors produced by the reference itself (tests/golden/make_golden_v3.py)."""
import json
import os

import numpy as np
import pytest

from helpers import INT32

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_v3.npz"))
CASES = json.load(open(os.path.join(HERE, "golden", "golden_v3.json")))["cases"]
NA32 = -2**31

SLICES = [(None, 2, None), (1, None, None), (None, None, 3), (2, 9, 3), (-3, None, None), (None, -2, None), (-4, -1, 2),
          (None, None, -1), (None, None, -2), (5, None, -1), (-2, None, -1), (5, 1, -2), (None, 2, -1), (-1, -5, -1),
          (10**6, None, None), (None, 0, None), (1, 3, 0), (-1, 2, 0), (-40, None, None), (None, None, -100), (3, 3, None),
          (0, 1, None), (-1, None, None), (None, None, 2**31 - 1), (None, None, -2**31)]


@pytest.mark.parametrize("device", [True, False])
def test_slice_groups_vs_oracle(device):
    import torch
    from datatable_b200 import engine
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    for ng, maxsize in ((1, 50), (7, 5), (1000, 40), (20_000, 9), (3, 100_000)):
        sizes = rng.integers(1, maxsize + 1, ng)
        offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
        off_in = torch.from_numpy(offsets).cuda() if device else offsets
        for s in SLICES:
            want_r, want_o = orc.slice_groups(offsets, *s)
            got_r, got_o = engine.slice_groups(off_in, *s)
            got_r = got_r.cpu().numpy() if device else got_r
            got_o = got_o.cpu().numpy() if device else got_o
            assert np.array_equal(got_o, want_o), (ng, s)
            assert np.array_equal(got_r, want_r), (ng, s)
    got_r, got_o = engine.slice_groups(np.zeros(1, np.int32), 1, 5, 2)           # no groups at all
    assert len(got_r) == 0 and np.array_equal(got_o, [0])


def test_slice_groups_refuses_bad_slices():
    from datatable_b200 import engine, _lib
    off = np.array([0, 3, 9], dtype=np.int32)
    with pytest.raises(_lib.DtbValueError):
        engine.slice_groups(off, None, 3, 0)              # repeat slice without a start
    with pytest.raises(_lib.DtbValueError):
        engine.slice_groups(off, 1, None, 2**40)          # step beyond int32


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_frame_i_by_matches_reference(case):
    import datatable_b200 as dt
    f, by, sort = dt.f, dt.by, dt.sort
    k, v = G[case["name"] + ".k"], G[case["name"] + ".v"]
    DT = dt.Frame(k=k, r=np.arange(len(k), dtype=np.int32), v=v)
    i = case["i"] if isinstance(case["i"], int) else slice(*case["i"])
    if case["mode"] == "by":
        R = DT[i, :, by(f.k)]
    elif case["mode"] == "sort":
        R = DT[i, :, sort(f.k)]
    else:
        R = DT[i, {"s": dt.sum(f.v), "n": dt.count(f.v), "first_r": dt.first(f.r)}, by(f.k)]
    assert R.nrows == case["nrows"]
    assert list(R.names) == case["names"]
    assert np.array_equal(R.to_numpy("k"), G[case["name"] + ".out_k"])
    if case["mode"] == "red":
        assert np.allclose(R.to_numpy("s"), G[case["name"] + ".out_s"], rtol=1e-12, atol=1e-12)
        assert np.array_equal(R.to_numpy("n"), G[case["name"] + ".out_n"])
        if case["nrows"]:
            assert np.array_equal(R.to_numpy("first_r"), G[case["name"] + ".out_first_r"])
    else:
        assert np.array_equal(R.to_numpy("r"), G[case["name"] + ".out_r"])


def test_frame_i_by_large_vs_oracle():
    import datatable_b200 as dt
    from oracle import oracle as orc
    f, by = dt.f, dt.by
    rng = np.random.default_rng(4)
    n = 1_000_003
    k = rng.integers(0, 50_000, n).astype(np.int32); k[::31] = NA32
    DT = dt.Frame(k=k, r=np.arange(n, dtype=np.int32))
    order, offsets, _ = orc.group([k], [0], 1, stypes=[INT32])
    for s in ((None, 3, None), (-2, None, None), (None, None, -1), (1, None, 4), 0, -1, 7):
        R = DT[s if isinstance(s, int) else slice(*s), :, by(f.k)]
        sel, off2 = orc.int_groups(offsets, s) if isinstance(s, int) else orc.slice_groups(offsets, *s)
        assert np.array_equal(R.to_numpy("r"), order[sel].astype(np.int32)), s


def test_frame_plain_row_slices():
    import datatable_b200 as dt
    DT = dt.Frame(A=list(range(10)), B=[x * 0.5 for x in range(10)])
    assert DT[2:9:3, :].to_list() == [[2, 5, 8], [1.0, 2.5, 4.0]]
    assert DT[::-4, :].to_list() == [[9, 5, 1], [4.5, 2.5, 0.5]]
    assert DT[-1, :].to_list() == [[9], [4.5]]
    assert DT[None, :].to_list() == DT[:, :].to_list()          # None selects every row (fexpr_literal_none.cc:88-96)
    with pytest.raises(ValueError):
        DT[10, :]
