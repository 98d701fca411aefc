"""CPU: pins the oracle (oracle/dt_oracle.c) against vectors produced by the reference itself."""
import numpy as np
import pytest

from conftest import golden
from helpers import NA_POS, OPS, case_flags, assert_reducer_equal
from oracle import oracle as orc

CASES = golden().cases


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_reference(case):
    g = golden()
    keys = [g.get(case, f"k{i}") for i in range(len(case["kst"]))]
    order, offsets, ng = orc.group(keys, case_flags(case), NA_POS[case["na_position"]], stypes=case["kst"])
    assert np.array_equal(order, g.get(case, "order")), "RowIndex differs from the reference"
    if case["nby"] is not None:
        want = g.get(case, "offsets")
        assert np.array_equal(offsets, want), "Groupby offsets differ from the reference"
        assert ng == len(want) - 1
        for j, (op, vi) in enumerate(case["reducers"]):
            v = g.get(case, f"v{vi}")
            got = orc.reduce(OPS[op], v, order, offsets, stype=case["vst"][vi])
            want_r = g.get(case, f"red{j}")
            if want_r.dtype != got.dtype and op in ("min", "max") and case["vst"][vi] == 1:
                want_r = want_r.astype(np.int8)
            assert_reducer_equal(got, want_r, op, case["vst"][vi], ctx=f"{op}(v{vi})")
    else:
        assert offsets is None


def test_gather_na():
    src = np.arange(10, dtype=np.float64)
    idx = np.array([3, -1, 0, 9, -2147483648], dtype=np.int32)
    out = orc.gather(src, idx)
    assert np.array_equal(np.isnan(out), [False, True, False, False, True])
    assert out[0] == 3 and out[2] == 0 and out[3] == 9
    srci = np.arange(10, dtype=np.int16)
    outi = orc.gather(srci, idx)
    assert outi.tolist() == [3, -32768, 0, 9, -32768]


def test_oracle_threads_do_not_change_results():
    """bench.py's CPU legs run the oracle on all host cores (orc_set_threads); every parallel region is a
    static partition with a deterministic combine, so the answers must be identical to one thread."""
    from oracle import oracle as orc
    rng = np.random.default_rng(11)
    n = 300_000
    k1 = rng.integers(-5000, 5000, n).astype(np.int32)
    k1[rng.random(n) < 0.01] = -2**31
    k2 = (rng.standard_normal(n) * 3).round(1)
    k2[rng.random(n) < 0.01] = np.nan
    v = rng.random(n)
    v[rng.random(n) < 0.05] = np.nan
    vi = rng.integers(-100, 100, n).astype(np.int16)
    cases = [([k1], [0], 1), ([k1, k2], [0, 2], 2), ([k2], [2], 3), ([k1, k2], [0, 4], 1)]
    try:
        for cols, flags, na_pos in cases:
            orc.set_threads(1)
            o1, f1, g1 = orc.group(cols, flags, na_pos)
            ops = (orc.SUM, orc.MEAN, orc.MIN, orc.MAX, orc.COUNT) if (f1 is not None and na_pos != 3) else ()
            r1 = [orc.reduce(op, val, o1, f1) for op in ops for val in (v, vi)]
            for t in (3, 8):
                orc.set_threads(t)
                assert orc.get_threads() == t
                o2, f2, g2 = orc.group(cols, flags, na_pos)
                assert g1 == g2 and np.array_equal(o1, o2)
                assert (f1 is None and f2 is None) or np.array_equal(f1, f2)
                r2 = [orc.reduce(op, val, o2, f2) for op in ops for val in (v, vi)]
                for a, b in zip(r1, r2):
                    assert np.array_equal(a, b, equal_nan=True)
    finally:
        orc.set_threads(1)
