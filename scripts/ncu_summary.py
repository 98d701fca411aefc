"""Summarise an .ncu-rep: per-kernel duration, DRAM traffic, occupancy, issue rate, top stall reasons,
and (optionally) the most-stalled SASS lines.  Usage: ncu_summary.py report.ncu-rep [--source N]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
def col(r, name):
    return r[hdr.index(name)] if name in hdr else ""
stall = [i for i, h in enumerate(hdr) if "smsp__average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = col(r, "Kernel Name")
    print("==", name[:110])
    rd, wr = float(col(r, "dram__bytes_read.sum") or 0), float(col(r, "dram__bytes_write.sum") or 0)
    u_rd = units[hdr.index("dram__bytes_read.sum")]
    print(f"   time {col(r,'gpu__time_duration.sum')} {units[hdr.index('gpu__time_duration.sum')]}  grid {col(r,'launch__grid_size')}  "
          f"regs {col(r,'launch__registers_per_thread')}  dram rd {rd:.3f} wr {wr:.3f} {u_rd}  "
          f"dram% {float(col(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed') or 0):.1f}  "
          f"warps_active% {float(col(r,'sm__warps_active.avg.pct_of_peak_sustained_active') or 0):.1f}  "
          f"issue% {float(col(r,'smsp__issue_active.avg.pct_of_peak_sustained_active') or 0):.1f}  "
          f"inst {float(col(r,'smsp__inst_executed.sum') or 0)/1e9:.2f}G  "
          f"smem_wavefronts {float(col(r,'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') or 0)/1e9:.2f}G  "
          f"bank_conflicts {float(col(r,'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') or 0)/1e9:.2f}G")
    vals = sorted([(float(r[i]) if r[i] else 0.0, hdr[i].split("stalled_")[1].split("_per_issue")[0]) for i in stall], reverse=True)
    print("   stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in vals[:7]))
if "--source" in sys.argv:
    topn = int(sys.argv[sys.argv.index("--source") + 1])
    kfilter = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else None
    args = ["ncu", "-i", rep, "--page", "source", "--csv"]
    if kfilter:
        args += ["--kernel-name", f"regex:{kfilter}"]
    skip = sys.argv[sys.argv.index("--skip") + 1] if "--skip" in sys.argv else "0"
    args += ["--launch-skip", skip, "--launch-count", "1"]
    src = subprocess.run(args, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    h = None; data = []; seen = set()
    for r in rows:
        if "Source" in r and "Address" in r:
            h = r; continue
        if h is None or len(r) != len(h):
            continue
        try:
            key = (r[h.index("Address")])
            if key in seen: continue
            seen.add(key)
            data.append((int(r[h.index("Warp Stall Sampling (All Samples)")]), int(r[h.index("Instructions Executed")]), r[h.index("Source")].strip()))
        except Exception:
            pass
    tot = sum(d[0] for d in data) or 1
    print(f"-- source: {len(data)} SASS lines, {tot} samples")
    for s_, e, t in sorted(data, reverse=True)[:topn]:
        print(f"   {100*s_/tot:5.1f}%  exec={e:>10d}  {t[:110]}")
