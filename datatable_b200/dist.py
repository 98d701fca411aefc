"""
Row-partitioned groupby over several GPUs of one box (one process per GPU,
torch.distributed / NCCL over NVLink).  The reference is single-process
(SURVEY.md 8e); this is the "final NCCL reduce of per-group partials" leg of
the north star: every rank groups and reduces its own row partition with the
single-GPU kernels, the per-group partials (key, partial) are exchanged with
one NCCL all-gather, and the same group()/reduce() kernels merge them.  Sums,
counts, minima and maxima are associative, so the merged result equals the
single-GPU result on the concatenated rows (float sums up to association).

No row ever crosses NVLink: the exchange moves 12-16 bytes per *group*.
The global RowIndex of a partitioned frame needs the radix-bucket all-to-all
(north star, config 5) and is not built yet.
"""
import torch
import torch.distributed as dist

from . import _lib, engine

_MERGE_OP = {_lib.OP_SUM: _lib.OP_SUM, _lib.OP_MIN: _lib.OP_MIN, _lib.OP_MAX: _lib.OP_MAX,
             _lib.OP_COUNT: _lib.OP_SUM, _lib.OP_NROWS: _lib.OP_SUM, _lib.OP_COUNTNA: _lib.OP_SUM}


def local_groupby(k, v, op):
    """(group keys, partials) of this rank's partition; k, v are CUDA tensors."""
    order, offsets, ng = engine.group([k], [0], _lib.NA_FIRST)
    part = engine.reduce(op, v, order, offsets)
    first = engine.gather(engine.Col(order, _lib.INT32), offsets[:-1])
    gkeys = engine.gather(k, first)
    return gkeys, part


class _EngineKernels:
    """The product path: libdtb200.so kernels.  (tests/ swap in an oracle-backed object to run the
    exchange logic under gloo on CPU.)"""
    group = staticmethod(lambda keys: engine.group([keys], [0], _lib.NA_FIRST))
    reduce = staticmethod(lambda op, v, order, offsets: engine.reduce(op, v, order, offsets))
    take = staticmethod(lambda src, idx: engine.gather(src, idx))


def merge_partials(gkeys, part, op, group=None, kernels=_EngineKernels):
    """All-gather every rank's (key, partial) list and merge equal keys with the engine's kernels."""
    world = dist.get_world_size(group)
    if world == 1:
        return gkeys, part
    n_local = torch.tensor([gkeys.numel()], dtype=torch.int64, device=gkeys.device)
    sizes = torch.empty(world, dtype=torch.int64, device=gkeys.device)
    dist.all_gather_into_tensor(sizes, n_local, group=group)
    sizes_h = sizes.tolist()
    cap = max(sizes_h)
    kpad = torch.zeros(cap, dtype=gkeys.dtype, device=gkeys.device); kpad[:gkeys.numel()] = gkeys
    ppad = torch.zeros(cap, dtype=part.dtype, device=part.device); ppad[:part.numel()] = part
    kall = torch.empty(world * cap, dtype=gkeys.dtype, device=gkeys.device)
    pall = torch.empty(world * cap, dtype=part.dtype, device=part.device)
    dist.all_gather_into_tensor(kall, kpad, group=group)
    dist.all_gather_into_tensor(pall, ppad, group=group)
    if any(sz != cap for sz in sizes_h):
        keep = torch.cat([torch.arange(r * cap, r * cap + sizes_h[r], device=gkeys.device) for r in range(world)])
        kall, pall = kall[keep].contiguous(), pall[keep].contiguous()
    order, offsets, ng = kernels.group(kall)
    merged = kernels.reduce(_MERGE_OP[op], pall, order, offsets)
    first = kernels.take(order, offsets[:-1])
    return kernels.take(kall, first), merged


def groupby_partitioned(k, v, op=_lib.OP_SUM, group=None):
    """DT[:, op(f.v), by(f.k)] over a frame row-partitioned across the ranks of `group`."""
    gkeys, part = local_groupby(k, v, op)
    if dist.is_available() and dist.is_initialized():
        return merge_partials(gkeys, part, op, group)
    return gkeys, part
