"""torchrun --nproc-per-node N scripts/dist_check.py [rows_per_rank]
Checks datatable_b200.dist on NCCL against the single-GPU engine on the concatenated data, then
times the C5-shaped paths (int64 keys, many groups) -- diagnostic, not a bench line."""
import os, sys, time, json
import torch, torch.distributed as dist
sys.path.insert(0, ".")
rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); lr = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from datatable_b200 import engine, _lib, dist as ddist

def gather_var(t):
    n = torch.tensor([t.numel()], device="cuda"); sizes = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(world)]
    dist.all_gather(sizes, n); sizes = [int(s) for s in sizes]
    bufs = [torch.empty(s, dtype=t.dtype, device="cuda") for s in sizes]
    dist.all_gather(bufs, t.contiguous())
    return torch.cat(bufs)

# ---------------- correctness at small size ----------------
n = 1_500_000 + 1000 * rank
g = torch.Generator(device="cuda"); g.manual_seed(100 + rank)
k = torch.randint(-2**40, 2**40, (n,), generator=g, device="cuda", dtype=torch.int64)
k[::7] = k[0]
kk = torch.randint(0, 50_000, (n,), generator=g, device="cuda", dtype=torch.int64)
v = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
row0 = sum(1_500_000 + 1000 * r for r in range(rank))
kall, kkall, vall = gather_var(k), gather_var(kk), gather_var(v)
ok = True
# groupby: allgather and alltoall variants vs single GPU
o, f, ng = engine.group([kkall], [0], 1)
want_s = engine.reduce(_lib.OP_SUM, vall, o, f)
want_k = engine.gather(kkall, engine.gather(engine.Col(o, _lib.INT32), f[:-1]))
gk, gs = ddist.groupby_partitioned(kk, v, _lib.OP_SUM)
ok &= bool(torch.equal(gk, want_k)) and bool(torch.allclose(gs, want_s, rtol=1e-9))
ak, as_ = ddist.groupby_partitioned(kk, v, _lib.OP_SUM, exchange="alltoall")
ak, as_ = gather_var(ak), gather_var(as_)
ok &= bool(torch.equal(ak, want_k)) and bool(torch.allclose(as_, want_s, rtol=1e-9))
# distributed sort vs single-GPU stable order (global row ids)
sk, sid = ddist.sort_partitioned(k, row0)
sid_all = gather_var(sid)
want_o = engine.group([kall], [_lib.FLAG_SORT_ONLY], 1)[0].long()
ok &= bool(torch.equal(sid_all, want_o))
res = {"world": world, "correct": bool(ok)}

# ---------------- timings at scale ----------------
rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 500_000_000
del kall, kkall, vall, k, kk, v
torch.cuda.empty_cache()
k = torch.randint(0, 100_000_000, (rows,), generator=g, device="cuda", dtype=torch.int64)   # C5: 1e8 distinct int64 keys
v = torch.rand(rows, generator=g, device="cuda", dtype=torch.float64)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); dist.barrier()
    return (time.perf_counter() - t0) / reps * 1e3
ms = timed(lambda: ddist.groupby_partitioned(k, v, _lib.OP_SUM, exchange="alltoall"))
res["C5_groupby_sum_int64_1e8keys_alltoall"] = {"rows_per_rank": rows, "ms": ms, "rows_per_s_total": world * rows / ms * 1e3}
ms = timed(lambda: ddist.sort_partitioned(k, rank * rows))
res["sort_partitioned_int64"] = {"rows_per_rank": rows, "ms": ms, "rows_per_s_total": world * rows / ms * 1e3}
if rank == 0:
    print(json.dumps(res, indent=1))
dist.destroy_process_group()
