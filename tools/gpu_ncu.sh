#!/bin/bash
# ncu passes (B200_PROFILING.md): launch list of one short step, then full captures of the top kernels.
# The cluster recurrences are launched without the cooperative attribute under ncu (NABU_REC_NOCOOP=1): ncu fails the
# launch of a cooperative cluster kernel (LaunchFailed before the kernel starts).
mkdir -p gpurun_out
export NABU_BENCH_T=${NABU_BENCH_T:-96}
export NABU_REC_NOCOOP=1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/launches_T${NABU_BENCH_T}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "launch list: exit $? lines $(wc -l < gpurun_out/launches_T${NABU_BENCH_T}.csv)"; grep blstm_rec gpurun_out/launches_T${NABU_BENCH_T}.csv | head -2 | cut -c1-400
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blstm_rec" -s 10 -c 2 -o gpurun_out/full_rec -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_rec.log 2>&1
tail -4 gpurun_out/ncu_full_rec.log | cut -c1-300
