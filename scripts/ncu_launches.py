"""Summarises an `ncu --csv --metrics ...` launch list: one line per launch (time, DRAM bytes, instructions,
LSU wavefronts, L2 atomics) and totals per kernel family.  Usage: ncu_launches.py file.csv [rows]"""
import collections
import csv
import sys

path = sys.argv[1]
rows_n = float(sys.argv[2]) if len(sys.argv) > 2 else None
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
hdr, rows = rows[0], rows[1:]
ix = {h: i for i, h in enumerate(hdr)}
by = collections.OrderedDict()
for r in rows:
    by.setdefault((int(r[ix['ID']]), r[ix['Kernel Name']]), {})[r[ix['Metric Name']]] = (r[ix['Metric Value']], r[ix['Metric Unit']])
UNIT = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-6, 'us': 1e-3, 'ms': 1, 's': 1e3}


def val(m, k):
    if k not in m:
        return 0.0
    v, u = m[k]
    return float(v.replace(',', '')) * UNIT.get(u, 1)


fam = collections.OrderedDict()
print(f"{'id':>3} {'kernel':42s} {'ms':>8} {'rd GB':>7} {'wr GB':>7} {'Minst':>8} {'issue%':>6} {'LSUwf M':>8} {'smem M':>8} {'warps%':>6} {'L2 atom/red M':>13}")
for (i, name), m in by.items():
    short = name.split('(')[0].replace('void ', '').replace('dtb::', '')
    short = short[:short.index('<')] if '<' in short else short
    t = val(m, 'gpu__time_duration.sum'); rd = val(m, 'dram__bytes_read.sum') / 1e9; wr = val(m, 'dram__bytes_write.sum') / 1e9
    inst = val(m, 'smsp__inst_executed.sum') / 1e6; lsu = val(m, 'l1tex__data_pipe_lsu_wavefronts.sum') / 1e6
    sm = val(m, 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum') / 1e6
    at = (val(m, 'lts__t_sectors_srcunit_tex_op_atom.sum') + val(m, 'lts__t_sectors_srcunit_tex_op_red.sum')) / 1e6
    print(f"{i:3d} {short:42s} {t:8.3f} {rd:7.2f} {wr:7.2f} {inst:8.1f} {val(m, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):6.1f} "
          f"{lsu:8.1f} {sm:8.1f} {val(m, 'sm__warps_active.avg.pct_of_peak_sustained_active'):6.1f} {at:13.1f}")
    f = fam.setdefault(short, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    f[0] += 1; f[1] += t; f[2] += rd + wr; f[3] += inst; f[4] += lsu; f[5] += at
tt = sum(f[1] for f in fam.values()); tb = sum(f[2] for f in fam.values())
print(f"\nper kernel family: launches, ms, share, DRAM GB, Minst, LSU wavefronts M, L2 atomics M" + (f"   [per row: n = {rows_n:.3g}]" if rows_n else ""))
for k, f in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    extra = f"   inst/row {f[3]*32e6/rows_n/f[0]:.1f}  LSUwf/row {f[4]*1e6/rows_n/f[0]:.2f} (per launch)" if rows_n and f[1] > 0.2 else ""
    print(f"  {k:42s} {f[0]:3d} {f[1]:8.3f} {100*f[1]/tt:5.1f}% {f[2]:8.2f} {f[3]:9.1f} {f[4]:9.1f} {f[5]:9.1f}{extra}")
print(f"  {'total':42s} {sum(f[0] for f in fam.values()):3d} {tt:8.3f}        {tb:8.2f}")
