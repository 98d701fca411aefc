#!/usr/bin/env python
"""
bench.py -- BASELINE.json's metric on BASELINE.json's config.

    metric  : rows/sec of DT[:, sum(f.v), by(f.k)]  (groupby-sum)
    config  : C2 = 1e9 rows, int32 key with 1e6 distinct values, float64 value, 1 x B200
              (N > 1: every rank owns its own 1e9-row partition -> weak scaling; per-group
               partials are merged with one NCCL all-gather + the same kernels)

One "step" = one pass of the hot path over one batch:
    group() -> RowIndex + Groupby offsets, then the per-group SUM reducer.

    value    device-resident inputs, CUDA-event timed, max over ranks
    e2e      the public Frame API on pinned HOST columns: H2D of k and v, the query,
             D2H of the result frame, all inside the timed region
    roofline dominant kernel (radix scatter pass): algorithmic bytes / CUDA-event time
    cpu_baseline   the CPU oracle port (or the reference build under oracle/_ref when
             present) on a bounded sample, timed on this box's host cores

`--impl reference` times the CPU implementation alone and prints the same line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "rows/sec groupby-sum 1e9 int32 keys"
UNIT = "rows/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000_000, help="rows per GPU (C2 = 1e9)")
    ap.add_argument("--groups", type=int, default=1_000_000)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=50_000_000, help="bounded CPU sample per step")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# clocks: nvidia-smi sampled DURING the timed region
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2]))
            except ValueError:
                continue
            for nm, val in zip(names, p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU side: the oracle port, or the reference itself when oracle/_ref holds a build of it
# ---------------------------------------------------------------------------
def cpu_groupby_sum(k, v):
    """Returns (seconds, kind, cores) for DT[:, sum(v), by(k)] on host arrays."""
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    if os.path.exists(os.path.join(ref_dir, "datatable", "__init__.py")):
        sys.path.insert(0, ref_dir)
        import datatable as rdt                         # the reference, built from /root/reference
        DT = rdt.Frame(k=k, v=v)
        t0 = time.perf_counter()
        R = DT[:, rdt.sum(rdt.f.v), rdt.by(rdt.f.k)]
        R.materialize()
        return time.perf_counter() - t0, "reference", int(rdt.options.nthreads)
    from oracle import oracle as orc
    orc.build()
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    cores = max(1, min(cores, 256))
    orc.set_threads(cores)                              # chunk-parallel like the reference's own sort
    try:
        t0 = time.perf_counter()
        o, f, ng = orc.group([k], [0], orc.NA_FIRST)
        orc.reduce(orc.SUM, v, o, f)
        dt = time.perf_counter() - t0
    finally:
        orc.set_threads(1)
    return dt, "port", cores


def host_sample(rows, groups, seed):
    rng = np.random.default_rng(seed)
    k = rng.integers(0, groups, rows, dtype=np.int32)
    v = rng.random(rows)
    return k, v


def run_reference(args, rank, world):
    if rank != 0:
        return
    rows = min(args.rows, args.cpu_rows)
    k, v = host_sample(rows, args.groups, 42)
    kind, cores = "port", 1
    for _ in range(max(0, min(args.warmup, 1))):
        cpu_groupby_sum(k, v)
    ts = []
    for _ in range(args.steps):
        t, kind, cores = cpu_groupby_sum(k, v)
        ts.append(t)
    ms = 1e3 * sum(ts) / len(ts)
    value = rows / (ms / 1e3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32 keys / float64 sums", "data": "synthetic",
        "config": {"workload": "C2: int32 key (1e6 distinct), float64 value, DT[:, sum(v), by(k)]",
                   "rows_per_step": rows, "groups": args.groups},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{rows} rows of the C2 workload per step (uniform keys in [0,{args.groups}))"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# the B200 arm
# ---------------------------------------------------------------------------
def run_b200(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (datatable_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import datatable_b200 as dtb
    from datatable_b200 import engine, _lib, dist as ddist
    from datatable_b200 import f, by

    n, G = args.rows, args.groups
    gen = torch.Generator(device="cuda"); gen.manual_seed(42 + rank)
    k = torch.randint(0, G, (n,), generator=gen, device="cuda", dtype=torch.int32)
    v = torch.rand(n, generator=gen, device="cuda", dtype=torch.float64)
    torch.cuda.synchronize()

    launches = [0]

    def step():
        # group(): RowIndex int32[n] + Groupby offsets int32[ng+1], both left in HBM behind the handle
        # and the SUM reducer, evaluated inside the same engine call (overlapped with the sort passes)
        gb = engine.Groupby([k], [0], _lib.NA_FIRST, reducers=[(_lib.OP_SUM, v)])
        launches[0] += _lib.last_call_stats()["kernels_launched"]
        sums = gb.reduced(0)
        ng = gb.ngroups
        if world > 1:
            gkeys = engine.gather(k, gb.first_rows())
            launches[0] += 2
            gkeys, sums = ddist.merge_partials(gkeys, sums, _lib.OP_SUM)
        gb.close()
        return None, None, ng, sums

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()

    engine.set_option("profile", 1)
    _lib.profile_records(reset=True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches[0] = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        out = step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    engine.set_option("profile", 0)
    prof = _lib.profile_records(reset=True)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * n / (ms_step / 1e3)
    ngroups = out[2]
    launches_per_step = launches[0] / args.steps

    # ---- sanity: the sums are the column total (cheap size-independent check, not timed) ----
    tot = float(out[3].sum().item())
    vsum = v.sum()
    if world > 1:
        dist.all_reduce(vsum)                   # the merged sums cover every rank's partition
    ref_tot = float(vsum.item())
    if abs(tot - ref_tot) > 1e-6 * abs(ref_tot):
        raise SystemExit(f"bench.py: group sums do not add up: {tot} vs {ref_tot}")

    # ---- roofline of the dominant kernel: the radix scatter passes --------------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    fam = {}
    for name, ms in prof:
        fam.setdefault(name, []).append(ms)
    passes = fam.get("radix_scatter", [])
    npass_step = len(passes) // max(1, args.steps)
    # algorithmic bytes of one scatter launch over n rows (32-bit keys, int32 row ids):
    #   first pass : read key 4 (row id = position) + write (key 4 + idx 4)  = 12 B/row
    #   middle pass: read (key 4 + idx 4)           + write (key 4 + idx 4)  = 16 B/row
    #   last pass  : read (key 4 + idx 4)           + write idx 4            = 12 B/row
    #                (small key domain: group sizes go to the count table, the sorted keys are not written)
    pass_bytes = ([12.0 * n] + [16.0 * n] * max(0, npass_step - 2) + [12.0 * n]) if npass_step >= 2 else [8.0 * n]
    alg_bytes_launch = sum(pass_bytes) / max(1, npass_step)
    pass_ms = sum(passes) / max(1, len(passes))
    achieved = alg_bytes_launch / (pass_ms / 1e3) / 1e9 if passes else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "scatter_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    roofline = {"kernel": "scatter_kernel (radix pass)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "peak_source": peak_src, "launches_per_step": npass_step, "avg_launch_ms": pass_ms,
                "alg_bytes_per_launch": alg_bytes_launch}
    step_alg = 16.0 * n / (ms_step / 1e3) / 1e9          # key 4 + value 8 + RowIndex 4 B/row
    kernel_ms = {nm: sum(v_) / args.steps for nm, v_ in fam.items()}
    if rank == 0 and os.environ.get("DTB_BENCH_DEBUG"):
        print("kernel_ms_per_step", kernel_ms, "step", ms_step, file=sys.stderr)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32 keys / float64 sums", "data": "synthetic",
        "config": {"workload": "C2: int32 key (1e6 distinct), float64 value, DT[:, sum(v), by(k)]",
                   "rows_per_gpu": n, "groups": G, "ngroups_found": ngroups,
                   "outputs": "RowIndex int32[n] + Groupby offsets int32[ng+1] + float64 sums[ng]",
                   "l2": "inputs (12 GB/GPU) exceed L2 (126 MB); no flush needed",
                   "parallelism": f"row-partitioned x{world}, NCCL all-gather of per-group partials" if world > 1 else "single GPU"},
        "roofline": roofline,
        "roofline_step": {"alg_bytes_per_row": 16, "achieved": step_alg, "peak": peak, "unit": "GB/s", "frac": step_alg / peak},
        "kernel_ms_per_step": kernel_ms,
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clocks,
    }

    # ---- e2e: public Frame API on pinned host columns ---------------------------------------
    if not args.no_e2e:
        kh = torch.empty(n, dtype=torch.int32, pin_memory=True); kh.copy_(k)
        vh = torch.empty(n, dtype=torch.float64, pin_memory=True); vh.copy_(v)
        torch.cuda.synchronize()
        DT = dtb.Frame(k=kh, v=vh)

        def e2e_step():
            R = DT[:, dtb.sum(f.v), by(f.k)]
            if world > 1:                                   # merge the per-rank result frames over NCCL
                gk = torch.from_numpy(R.to_numpy("k")).cuda()
                gs = torch.from_numpy(R.to_numpy("v")).cuda()
                gk, gs = ddist.merge_partials(gk, gs, _lib.OP_SUM)
                gs.cpu()
            return R
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            R = e2e_step()
        torch.cuda.synchronize()
        dt_e2e = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt_e2e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt_e2e = float(t.item())
        e2e_ms = 1e3 * dt_e2e / args.e2e_steps
        line["e2e"] = {"value": world * n / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms,
                       "steps": args.e2e_steps,
                       "h2d_bytes_per_step": int(n * 12), "d2h_bytes_per_step": int(R.nrows * 12),
                       "api": "datatable_b200.Frame[:, sum(f.v), by(f.k)] on pinned host columns"}
        del kh, vh, DT

    # ---- CPU baseline on this box's host cores (rank 0, N = 1 only) ---------------------------
    if rank == 0 and world == 1 and not args.no_cpu:
        rows = min(n, args.cpu_rows)
        kc = k[:rows].cpu().numpy(); vc = v[:rows].cpu().numpy()
        secs, kind, cores = cpu_groupby_sum(kc, vc)
        line["cpu_baseline"] = {"value": rows / secs, "unit": UNIT, "cores": cores, "kind": kind,
                                "seconds": secs,
                                "sample": f"first {rows} rows of the same C2 input (keys uniform in [0,{G}))"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
