"""CPU: the oracle's restatement of the integer / slice `i` node under by() and sort()
(expr/fexpr_literal_sliceint.cc:82-170, fexpr_literal_int.cc:146-192, eval_context.cc:154-163) against vectors
produced by the reference itself (tests/golden/make_golden_v3.py)."""
import json
import os

import numpy as np
import pytest

from oracle import oracle as orc
from helpers import OPS

HERE = os.path.dirname(os.path.abspath(__file__))
G = np.load(os.path.join(HERE, "golden", "golden_v3.npz"))
CASES = json.load(open(os.path.join(HERE, "golden", "golden_v3.json")))["cases"]
INT32, FLOAT64 = 4, 7
SORT_ONLY = 4


def oracle_rows(case):
    """(original row numbers in output order, offsets of the remaining groups) as the oracle computes them"""
    k = G[case["name"] + ".k"]
    flags = [SORT_ONLY] if case["mode"] == "sort" else [0]
    order, offsets, _ = orc.group([k], flags, 1, stypes=[INT32])
    if offsets is None:
        offsets = np.array([0, len(k)], dtype=np.int32)          # sort(): one group (Groupby::single_group)
    i = case["i"]
    sel, off2 = orc.int_groups(offsets, i) if isinstance(i, int) else orc.slice_groups(offsets, *i)
    return order[sel], off2


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_iby_matches_reference(case):
    k, v = G[case["name"] + ".k"], G[case["name"] + ".v"]
    rows, off2 = oracle_rows(case)
    assert len(rows) == off2[-1]
    if case["mode"] in ("by", "sort"):
        assert case["nrows"] == len(rows)
        assert np.array_equal(G[case["name"] + ".out_r"], rows.astype(np.int32))
        assert np.array_equal(G[case["name"] + ".out_k"], k[rows])
    else:
        ng = len(off2) - 1
        assert case["nrows"] == ng
        ident = np.arange(len(rows), dtype=np.int32)
        assert np.array_equal(G[case["name"] + ".out_k"], k[rows[off2[:-1]]])
        if ng:       # (no groups left: the reference's first() yields one NA row in a frame of zero rows -- not restated)
            assert np.array_equal(G[case["name"] + ".out_first_r"], rows[off2[:-1]].astype(np.int32))
        s = orc.reduce(OPS["sum"], v[rows], ident, off2, stype=FLOAT64)
        assert np.allclose(G[case["name"] + ".out_s"], s, rtol=1e-12, atol=1e-12)
        assert np.array_equal(G[case["name"] + ".out_n"], np.diff(off2).astype(np.int64))    # v has no NA
