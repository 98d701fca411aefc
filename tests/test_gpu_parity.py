"""GPU: the CUDA path (through the C-ABI) against the reference's golden vectors and the oracle."""
import numpy as np
import pytest

from conftest import golden
from helpers import NA_POS, OPS, case_flags, assert_reducer_equal, BOOL

pytestmark = pytest.mark.gpu

CASES = golden().cases


def _run_case(case, device):
    import torch
    from datatable_b200 import engine
    g = golden()
    nk = len(case["kst"])
    keys_np = [g.get(case, f"k{i}") for i in range(nk)]
    if device:
        keys = [engine.Col(torch.from_numpy(k).cuda(), st) for k, st in zip(keys_np, case["kst"])]
    else:
        keys = [engine.Col(k, st) for k, st in zip(keys_np, case["kst"])]
    order, offsets, ng = engine.group(keys, case_flags(case), NA_POS[case["na_position"]])
    order_h = order.cpu().numpy() if device else order
    assert np.array_equal(order_h, g.get(case, "order")), "RowIndex differs from the reference"
    if case["nby"] is None:
        assert offsets is None
        return
    want = g.get(case, "offsets")
    offs_h = offsets.cpu().numpy() if device else offsets
    assert np.array_equal(offs_h, want), "Groupby offsets differ from the reference"
    assert ng == len(want) - 1
    for j, (op, vi) in enumerate(case["reducers"]):
        v = g.get(case, f"v{vi}")
        vst = case["vst"][vi]
        vv = engine.Col(torch.from_numpy(v).cuda(), vst) if device else engine.Col(v, vst)
        got = engine.reduce(OPS[op], vv, order, offsets)
        got = got.cpu().numpy() if device else got
        want_r = g.get(case, f"red{j}")
        if vst == BOOL and op in ("min", "max"):
            want_r = want_r.astype(np.int8)
        assert_reducer_equal(got, want_r, op, vst, ctx=f"{case['name']}:{op}(v{vi})")


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_golden_host_buffers(case):
    _run_case(case, device=False)


@pytest.mark.parametrize("case", CASES[::3], ids=[c["name"] for c in CASES[::3]])
def test_golden_device_buffers(case):
    _run_case(case, device=True)


def test_empty_and_single_row():
    from datatable_b200 import engine
    o, f, ng = engine.group([np.zeros(0, np.int32)])
    assert len(o) == 0 and f.tolist() == [0] and ng == 0           # sort.cc:1431-1434
    o, f, ng = engine.group([np.array([5], np.int64)])
    assert o.tolist() == [0] and f.tolist() == [0, 1] and ng == 1  # sort.cc:1435-1439
    o, f, ng = engine.group([np.array([5.0], np.float64)], [4])
    assert o.tolist() == [0] and f is None


def test_gather_na_index():
    from datatable_b200 import engine
    from oracle import oracle as orc
    rng = np.random.default_rng(3)
    for dt in (np.int8, np.int16, np.int32, np.int64, np.float32, np.float64):
        src = (rng.standard_normal(1000) * 100).astype(dt)
        idx = rng.integers(-1, 1000, 5000).astype(np.int32)
        idx[::7] = -2**31
        got = engine.gather(src, idx)
        want = orc.gather(src, idx)
        assert got.tobytes() == want.tobytes()
        got64 = engine.gather(src, idx.astype(np.int64))
        assert got64.tobytes() == want.tobytes()
