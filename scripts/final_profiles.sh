#!/bin/bash
# One GPU call that produces the round's ncu evidence under gpurun_out/ (copied / summarised into profiles/ afterwards):
#   launch lists (time, DRAM bytes, instructions, LSU wavefronts, L2 atomics per launch) of one C2 / C3 / C4 call at 1e9 rows,
#   the launch list of the bench command itself, and --set full captures of the scatter passes and the bucket kernels.
set -u
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_atom.sum,lts__t_sectors_srcunit_tex_op_red.sum
TAG=${1:-r3}
for c in c2 c3 c4; do
  ncu --metrics $M --clock-control none --profile-from-start off --csv --log-file gpurun_out/${TAG}_${c}_launches_1e9.csv \
      python scripts/prof_case.py $c 1e9 > gpurun_out/${TAG}_prof_${c}.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extra --no-e2e --no-cpu > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scatter_kernel \
    -o gpurun_out/${TAG}_c2_scatter_full -f python scripts/prof_case.py c2 1e9 > gpurun_out/${TAG}_prof_c2_full.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:bucket_ \
    -o gpurun_out/${TAG}_c4_bucket_full -f python scripts/prof_case.py c4 1e9 > gpurun_out/${TAG}_prof_c4_full.log 2>&1
ls -la gpurun_out/${TAG}_*
