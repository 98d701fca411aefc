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
