#!/bin/bash
# One GPU-box pass: parity tests, bench (both workloads), ncu launch list + one full capture.
# Outputs land in gpurun_out/ (merged back by gpurun).
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ctc.json 2> gpurun_out/bench_ctc.err
timeout 600 python bench.py --steps 5 --warmup 3 --workload las --no-cpu-baseline > gpurun_out/bench_las.json 2> gpurun_out/bench_las.err
if [ "${NCU:-1}" = "1" ]; then
NABU_BENCH_T=96 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches_T96.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
NABU_BENCH_T=96 timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:"${NCU_K:-rec_}" -s ${NCU_S:-20} -c ${NCU_C:-4} -o gpurun_out/full_T96 -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ctc.json; cat gpurun_out/bench_las.json
