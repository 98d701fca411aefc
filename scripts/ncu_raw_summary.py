"""Summarise an .ncu-rep (ncu --set full) per launch: duration, DRAM bytes, instructions, LSU wavefronts, occupancy,
issue utilisation and the stall-reason samples.  Usage: python scripts/ncu_raw_summary.py report.ncu-rep [rows]"""
import csv, subprocess, sys
rep = sys.argv[1]
rows_n = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]; idx = {h: i for i, h in enumerate(hdr)}
def g(r, k, d=0.0):
    try: return float(r[idx[k]])
    except Exception: return d
stalls = [h for h in hdr if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")]
for r in rows[2:]:
    name = r[idx["Kernel Name"]].split("(")[0]
    ms = g(r, "gpu__time_duration.sum")
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    inst = g(r, "smsp__inst_executed.sum")
    wf = g(r, "l1tex__data_pipe_lsu_wavefronts.sum")
    print(f"{name}")
    print(f"   {ms:8.3f} ms   DRAM read {rd:7.2f} GB  write {wr:7.2f} GB  -> {(rd + wr) / ms:6.2f} TB/s"
          + (f"   ({g(r, 'dram__throughput.avg.pct_of_peak_sustained_elapsed'):.0f} % of DRAM peak)" if 'dram__throughput.avg.pct_of_peak_sustained_elapsed' in idx else ""))
    per = f"   per row: {inst * 32 / rows_n:6.1f} thread-instructions" if rows_n else ""
    wfs = f"{wf / 1e6:9.1f} M " if wf else ""
    print(f"   warp instructions {inst / 1e6:9.1f} M   LSU data-pipe wavefronts {wfs}"
          f"({g(r, 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed'):.1f} % of peak; shared "
          f"{g(r, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6:.1f} M, of which bank conflicts "
          f"{g(r, 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum') / 1e6:.1f} M){per}")
    print(f"   registers {g(r, 'launch__registers_per_thread'):.0f}   CTAs/SM by registers {g(r, 'launch__occupancy_limit_registers'):.0f}"
          f" / shared memory {g(r, 'launch__occupancy_limit_shared_mem'):.0f}   warps active {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} %"
          f"   issue slots {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} %")
    st = sorted(((g(r, k), k.replace("smsp__pcsamp_warps_issue_stalled_", "")) for k in stalls), reverse=True)
    tot = sum(v for v, _ in st) or 1.0
    print("   stall samples: " + ", ".join(f"{k} {100 * v / tot:.0f} %" for v, k in st[:7]))
