#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list of one short step, then full captures of the top kernels.
# The cluster recurrences are launched without the cooperative attribute under ncu (NABU_REC_NOCOOP=1): ncu fails the
# launch of a cooperative cluster kernel (LaunchFailed before the kernel starts).
mkdir -p gpurun_out
export NABU_BENCH_T=${NABU_BENCH_T:-96}
export NABU_REC_NOCOOP=1
export NABU_OVERLAP=0
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_T${NABU_BENCH_T}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list: exit $? lines $(wc -l < gpurun_out/launches_T${NABU_BENCH_T}.csv)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blstm_rec_fwd" -s 5 -c 1 -o gpurun_out/full_rec_fwd -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_rec_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blstm_rec_bwd" -s 5 -c 1 -o gpurun_out/full_rec_bwd -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_rec_bwd.log 2>&1
tail -2 gpurun_out/ncu_full_rec_bwd.log | cut -c1-200
# full-length GEMMs: one capture of each mode at the real cfg-3 shapes
unset NABU_BENCH_T
timeout 900 ncu --set full --clock-control none -k regex:"gemm_h2" -s 30 -c 6 -o gpurun_out/full_gemm -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_gemm.log 2>&1
tail -2 gpurun_out/ncu_full_gemm.log | cut -c1-200
