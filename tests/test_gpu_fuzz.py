"""GPU: seeded random "attack" over the whole C-ABI surface (the reference has tests_random/ for this):
random number of keys, stypes, directions, by/sort split, NA position, sizes and value columns; every
draw is checked against the oracle bit-exactly (RowIndex, offsets, integer reducers) or to 1e-6."""
import numpy as np
import pytest

from helpers import (BOOL, INT8, INT16, INT32, INT64, FLOAT32, FLOAT64, DESCENDING, SORT_ONLY,
                     OPS, assert_reducer_equal)
from test_gpu_random import make_col

pytestmark = pytest.mark.gpu
STYPES = [BOOL, INT8, INT16, INT32, INT64, FLOAT32, FLOAT64]


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_group_reduce_vs_oracle(seed):
    import torch
    from datatable_b200 import engine
    from oracle import oracle as orc
    rng = np.random.default_rng(7000 + seed)
    n = int(rng.choice([1, 2, 3, 64, 65, 257, 4095, 4096, 4097, 20_000, 70_000, 131_073]))
    nk = int(rng.integers(1, 4))
    sts = [int(rng.choice(STYPES)) for _ in range(nk)]
    keys = [make_col(rng, st, n, str(rng.choice(["few", "few", "unit", "wide"])), float(rng.choice([0.0, 0.05, 0.5])))
            for st in sts]
    nby = int(rng.integers(0, nk + 1))
    flags = [(DESCENDING if rng.integers(0, 2) else 0) | (SORT_ONLY if i >= nby else 0) for i in range(nk)]
    na_pos = int(rng.choice([1, 2])) if nby else int(rng.choice([1, 2, 3]))
    device = bool(rng.integers(0, 2))
    want_o, want_f, want_ng = orc.group(keys, flags, na_pos, stypes=sts)
    cols = [engine.Col(torch.from_numpy(k).cuda() if device else k, st) for k, st in zip(keys, sts)]
    got_o, got_f, got_ng = engine.group(cols, flags, na_pos)
    tonp = (lambda t: t.cpu().numpy()) if device else (lambda t: t)
    assert np.array_equal(tonp(got_o), want_o), f"seed {seed}: RowIndex"
    if nby == 0:
        assert got_f is None
        return
    assert np.array_equal(tonp(got_f), want_f) and got_ng == want_ng, f"seed {seed}: offsets"
    vst = int(rng.choice(STYPES))
    v = make_col(rng, vst, n, "few" if vst == BOOL else "unit", 0.2)
    vv = engine.Col(torch.from_numpy(v).cuda() if device else v, vst)
    for op in ("sum", "mean", "min", "max", "count", "countna"):
        want = orc.reduce(OPS[op], v, want_o, want_f, stype=vst)
        got = tonp(engine.reduce(OPS[op], vv, got_o, got_f))
        assert_reducer_equal(got, want, op, vst, ctx=f"seed {seed} {op}")
    if device and na_pos != 3:
        # the same through the fused handle API
        gb = engine.Groupby(cols, flags, na_pos, reducers=[(OPS["sum"], vv), (OPS["max"], vv), (OPS["nrows"], None)])
        assert gb.ngroups == want_ng
        assert_reducer_equal(gb.reduced(0).cpu().numpy(), orc.reduce(OPS["sum"], v, want_o, want_f, stype=vst), "sum", vst, ctx=f"seed {seed} fused sum")
        assert_reducer_equal(gb.reduced(1).cpu().numpy(), orc.reduce(OPS["max"], v, want_o, want_f, stype=vst), "max", vst, ctx=f"seed {seed} fused max")
        assert np.array_equal(gb.reduced(2).cpu().numpy(), np.diff(want_f).astype(np.int64))
        gb.close()


def test_reduce_rejects_offsets_that_are_not_a_groupby():
    """dtb_reduce validates caller-supplied offsets (groupby.h:41-47: offsets[0] = 0, strictly increasing):
    an empty group would silently shift every later group of its tile."""
    import torch
    from datatable_b200 import engine, _lib
    v = np.arange(10, dtype=np.float64)
    order = np.arange(10, dtype=np.int32)
    for bad in ([0, 3, 3, 10], [1, 3, 10], [0, 7, 5, 10]):
        offs = np.array(bad, dtype=np.int32)
        with pytest.raises(_lib.DtbValueError, match="not a Groupby"):
            engine.reduce(_lib.OP_SUM, v, order, offs)
        with pytest.raises(_lib.DtbValueError, match="not a Groupby"):
            engine.reduce(_lib.OP_SUM, torch.from_numpy(v).cuda(), torch.from_numpy(order).cuda(), torch.from_numpy(offs).cuda())
    got = engine.reduce(_lib.OP_SUM, v, order, np.array([0, 3, 10], dtype=np.int32))
    assert got.tolist() == [3.0, 42.0]


def test_constant_by_column_with_full_width_sort_column():
    """by(constant column) + sort(32-/64-bit column): the by-columns contribute 0 bits, so group_shift would equal
    the key width (a shift by the full width is undefined): the engine must return ONE group and the plain sort."""
    from datatable_b200 import engine, _lib
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    n = 50_000
    c = np.full(n, 7, dtype=np.int32)
    for x in (rng.integers(-2**31 + 1, 2**31 - 1, n, dtype=np.int64).astype(np.int32),
              rng.integers(-2**63 + 1, 2**63 - 1, n, dtype=np.int64)):
        want_o, want_f, want_ng = orc.group([c, x], [0, orc.SORT_ONLY], orc.NA_FIRST)
        o, f, ng = engine.group([c, x], [0, _lib.FLAG_SORT_ONLY], _lib.NA_FIRST)
        assert ng == want_ng == 1 and np.array_equal(f, want_f) and np.array_equal(o, want_o)
        gb = engine.Groupby([__import__("torch").from_numpy(c).cuda(), __import__("torch").from_numpy(x).cuda()],
                            [0, _lib.FLAG_SORT_ONLY], _lib.NA_FIRST)
        assert gb.ngroups == 1 and np.array_equal(gb.order().cpu().numpy(), want_o)
        gb.close()


def test_one_thread_two_devices():
    """The scratch arena follows the calling thread's current device: alternate two GPUs."""
    import torch
    from datatable_b200 import engine, _lib
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(8)
    k = rng.integers(0, 1000, 300_000).astype(np.int32)
    want = np.argsort(k, kind="stable").astype(np.int32)
    try:
        for dev in (0, 1, 0, 1):
            torch.cuda.set_device(dev)
            kd = torch.from_numpy(k).to(f"cuda:{dev}")
            o, f, ng = engine.group([kd], [0], _lib.NA_FIRST)
            assert o.device.index == dev and np.array_equal(o.cpu().numpy(), want)
    finally:
        torch.cuda.set_device(0)
