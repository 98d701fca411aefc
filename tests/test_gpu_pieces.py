"""GPU: reducers fed piecewise (dtb_groupby_reduce_begin / _add / _end) -- what the Frame does with a host value
column that is still on its way over PCIe -- against the whole-column reducer and the oracle."""
import numpy as np
import pytest

from helpers import INT32, INT64, FLOAT32, FLOAT64, OPS, assert_reducer_equal

pytestmark = pytest.mark.gpu


def _cuts(n, rng, k):
    c = sorted(set([0, n] + [int(x) for x in rng.integers(1, n, k)]))
    return list(zip(c[:-1], c[1:]))


@pytest.mark.parametrize("vst", [INT32, INT64, FLOAT32, FLOAT64])
def test_reduce_pieces_equals_whole_column_and_oracle(vst):
    import torch
    from datatable_b200 import engine, _lib
    from oracle import oracle as orc
    rng = np.random.default_rng(5 + vst)
    n = 400_003
    k = rng.integers(0, 5000, n).astype(np.int32); k[::17] = np.iinfo(np.int32).min
    if vst in (FLOAT32, FLOAT64):
        v = (rng.random(n) + 0.5).astype(np.float32 if vst == FLOAT32 else np.float64); v[::13] = np.nan   # positive: sums do not cancel
    else:
        v = rng.integers(-1000, 1000, n).astype(np.int32 if vst == INT32 else np.int64); v[::13] = np.iinfo(v.dtype).min
    kd, vd = torch.from_numpy(k).cuda(), torch.from_numpy(v).cuda()
    gb = engine.Groupby([engine.Col(kd, INT32)], [0], _lib.NA_FIRST)
    order, offsets, _ = orc.group([k], [0], _lib.NA_FIRST, stypes=[INT32])
    try:
        for op in ("sum", "mean", "min", "max", "count", "countna"):
            pcs = [(vd[a:b], a, None) for a, b in _cuts(n, rng, 6)]
            got = gb.reduce_pieces(OPS[op], vst, pcs)
            assert got is not None, "a 5000-key int32 domain has the streaming path"
            whole = gb.reduce(OPS[op], engine.Col(vd, vst))
            want = orc.reduce(OPS[op], v, order, offsets, stype=vst)
            assert_reducer_equal(got.cpu().numpy(), want, op, vst, f"pieces op={op}")
            assert_reducer_equal(whole.cpu().numpy(), want, op, vst, f"whole op={op}")
        # rows missing / repeated: refused at _end, state freed
        with pytest.raises(_lib.DtbValueError):
            gb.reduce_pieces(_lib.OP_SUM, vst, [(vd[: n // 2], 0, None)])
        with pytest.raises(_lib.DtbValueError):
            gb.reduce_pieces(_lib.OP_SUM, vst, [(vd[: n // 2], n - 10, None)])
    finally:
        gb.close()


def test_reduce_pieces_declined_without_streaming_path():
    import torch
    from datatable_b200 import engine, _lib
    rng = np.random.default_rng(3)
    n = 50_000
    k = rng.integers(-2**62, 2**62, n).astype(np.int64)           # key domain far beyond 2^22: RowIndex path only
    v = torch.from_numpy(rng.standard_normal(n)).cuda()
    gb = engine.Groupby([engine.Col(torch.from_numpy(k).cuda(), INT64)], [0], _lib.NA_FIRST)
    try:
        assert gb.reduce_pieces(_lib.OP_SUM, FLOAT64, [(v, 0, None)]) is None
        assert gb.reduce_pieces(_lib.OP_FIRST, FLOAT64, [(v, 0, None)]) is None
    finally:
        gb.close()


def test_frame_uploads_and_reduces_host_value_column_in_pieces(monkeypatch):
    import datatable_b200 as dt
    from datatable_b200 import frame, engine
    f, by = dt.f, dt.by
    rng = np.random.default_rng(8)
    n = 300_000
    k = rng.integers(0, 1000, n).astype(np.int32)
    v = rng.random(n)
    calls = []
    real = engine.Groupby.reduce_pieces
    monkeypatch.setattr(engine.Groupby, "reduce_pieces", lambda self, op, st, pcs: calls.append(len(pcs)) or real(self, op, st, pcs))
    monkeypatch.setattr(frame, "_PIECE_BYTES", 1 << 18)          # 2.4 MB column -> 10 pieces
    R = dt.Frame(k=k, v=v)[:, {"s": dt.sum(f.v), "n": dt.count()}, by(f.k)]
    assert calls and calls[0] == (n * 8 + (1 << 18) - 1) // (1 << 18)
    assert np.array_equal(R.to_numpy("k"), np.arange(1000, dtype=np.int32))
    want = np.bincount(k, weights=v, minlength=1000)
    assert np.allclose(R.to_numpy("s"), want, rtol=1e-12)
    assert np.array_equal(R.to_numpy("n"), np.bincount(k, minlength=1000))


def test_repeated_host_frame_queries_reuse_their_staging_buffers():
    """Every query on a host frame uploads its columns into device buffers from torch's caching allocator; they must
    come from the same per-stream pool every time (a fresh copy stream per query leaked 12 GB per C2 query)."""
    import torch
    import datatable_b200 as dt
    f, by = dt.f, dt.by
    rng = np.random.default_rng(1)
    n = 4_000_000
    DT = dt.Frame(k=rng.integers(0, 1000, n).astype(np.int32), v=rng.random(n))
    reserved = []
    for _ in range(5):
        R = DT[:, dt.sum(f.v), by(f.k)]
        assert R.nrows == 1000
        del R
        torch.cuda.synchronize()
        reserved.append(torch.cuda.memory_reserved())
    assert reserved[-1] == reserved[1], reserved
