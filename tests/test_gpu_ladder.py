"""GPU parity ladder: the BASELINE config shapes C2, C3, C4 at 1e7 and 1e8 rows, CUDA path
through the C-ABI against the CPU oracle on the same seeded inputs.

Bar: RowIndex and Groupby offsets bit-exact; integer/count/min/max reducers bit-exact; float sums and
means within 1e-6 relative (north_star).  Input list follows SURVEY.md 8(d): C2 uniform / 1 % NA / Zipf,
C3 N(0,1) / U[0,1) / adversarial (+-0, +-inf, NaN, denormals), C4 keys (id << 33, int32 < 1000) with 1 % NaN
values and all 12 reducers.  (The reference's own known-answer vectors are tests/test_gpu_parity.py; the
full 1e9-row sizes are checked by size-independent properties in bench.py / tests/test_gpu_random.py.)
"""
import os

import numpy as np
import pytest

from helpers import assert_reducer_equal, FLOAT64

pytestmark = pytest.mark.gpu

SIZES = [10_000_000, 100_000_000]


def _orc():
    from oracle import oracle as orc
    orc.build()
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    orc.set_threads(max(1, min(cores, 256)))       # results do not depend on the thread count
    return orc


def _dev(a):
    import torch
    return torch.from_numpy(a).cuda()


def _c2_keys(rng, n, dist):
    if dist == "uniform":
        return rng.integers(0, 1_000_000, n, dtype=np.int32)
    if dist == "na1pct":
        k = rng.integers(0, 1_000_000, n, dtype=np.int32)
        k[rng.random(n) < 0.01] = np.iinfo(np.int32).min          # NA keys form the first group
        return k
    # Zipf-like: a heavy head and a long tail inside [0, 1e6)
    u = rng.random(n)
    return np.minimum((u ** 6 * 1_000_000).astype(np.int32), 999_999)


@pytest.mark.parametrize("dist", ["uniform", "na1pct", "zipf"])
@pytest.mark.parametrize("n", SIZES)
def test_c2_groupby_sum_vs_oracle(n, dist):
    from datatable_b200 import engine, _lib
    orc = _orc()
    try:
        rng = np.random.default_rng(1000 + n % 997 + len(dist))
        k = _c2_keys(rng, n, dist)
        v = rng.random(n)
        v[::1013] = np.nan
        want_o, want_f, want_ng = orc.group([k], [0], orc.NA_FIRST)
        want_s = orc.reduce(orc.SUM, v, want_o, want_f)
        gb = engine.Groupby([_dev(k)], [0], _lib.NA_FIRST, reducers=[(_lib.OP_SUM, _dev(v))])
        try:
            assert gb.ngroups == want_ng
            assert np.array_equal(gb.order().cpu().numpy(), want_o), f"C2 {dist} n={n}: RowIndex differs"
            assert np.array_equal(gb.offsets().cpu().numpy(), want_f), f"C2 {dist} n={n}: offsets differ"
            assert_reducer_equal(gb.reduced(0).cpu().numpy(), want_s, "sum", FLOAT64, f"C2 {dist} n={n}")
        finally:
            gb.close()
    finally:
        orc.set_threads(1)


def _c3_keys(rng, n, dist):
    if dist == "normal":
        return rng.standard_normal(n)
    if dist == "uniform":
        return rng.random(n)
    x = rng.standard_normal(n)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 5e-324, -5e-324, 2.2250738585072014e-308,
                        -2.2250738585072014e-308, 1.7976931348623157e308, -1.7976931348623157e308])
    pick = rng.random(n) < 0.2
    x[pick] = special[rng.integers(0, len(special), int(pick.sum()))]
    den = rng.random(n) < 0.05                                    # random denormals of both signs
    x[den] = (rng.integers(1, 1 << 52, int(den.sum())).astype(np.uint64)
              | (rng.integers(0, 2, int(den.sum())).astype(np.uint64) << np.uint64(63))).view(np.float64)
    return x


@pytest.mark.parametrize("dist", ["normal", "uniform", "adversarial"])
@pytest.mark.parametrize("n", SIZES)
def test_c3_float64_sort_vs_oracle(n, dist):
    from datatable_b200 import engine, _lib
    orc = _orc()
    try:
        rng = np.random.default_rng(2000 + n % 991 + len(dist))
        x = _c3_keys(rng, n, dist)
        want_o, _, _ = orc.group([x], [orc.SORT_ONLY], orc.NA_FIRST)
        gb = engine.Groupby([_dev(x)], [_lib.FLAG_SORT_ONLY], _lib.NA_FIRST)
        try:
            assert gb.ngroups == -1
            assert np.array_equal(gb.order().cpu().numpy(), want_o), f"C3 {dist} n={n}: RowIndex differs"
        finally:
            gb.close()
    finally:
        orc.set_threads(1)


@pytest.mark.parametrize("n", SIZES)
def test_c4_two_keys_twelve_reducers_vs_oracle(n):
    from datatable_b200 import engine, _lib
    orc = _orc()
    try:
        rng = np.random.default_rng(3000 + n % 983)
        k1 = rng.integers(0, 1000, n).astype(np.int64) << 33
        k2 = rng.integers(0, 1000, n).astype(np.int32)
        vs = []
        for _ in range(3):
            v = rng.standard_normal(n)
            v[rng.random(n) < 0.01] = np.nan
            vs.append(v)
        ops = [("mean", _lib.OP_MEAN, orc.MEAN), ("min", _lib.OP_MIN, orc.MIN), ("max", _lib.OP_MAX, orc.MAX),
               ("count", _lib.OP_COUNT, orc.COUNT)]
        want_o, want_f, want_ng = orc.group([k1, k2], [0, 0], orc.NA_FIRST)
        dk1, dk2, dvs = _dev(k1), _dev(k2), [_dev(v) for v in vs]
        gb = engine.Groupby([dk1, dk2], [0, 0], _lib.NA_FIRST,
                            reducers=[(op, dv) for dv in dvs for (_, op, _) in ops])
        try:
            assert gb.ngroups == want_ng
            assert np.array_equal(gb.order().cpu().numpy(), want_o), f"C4 n={n}: RowIndex differs"
            assert np.array_equal(gb.offsets().cpu().numpy(), want_f), f"C4 n={n}: offsets differ"
            i = 0
            for ci, v in enumerate(vs):
                for name, _, oop in ops:
                    want = orc.reduce(oop, v, want_o, want_f)
                    assert_reducer_equal(gb.reduced(i).cpu().numpy(), want, name, FLOAT64, f"C4 n={n} col {ci} {name}")
                    i += 1
        finally:
            gb.close()
    finally:
        orc.set_threads(1)
