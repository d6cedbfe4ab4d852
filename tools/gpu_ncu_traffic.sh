#!/bin/bash
# DRAM traffic of the recurrence kernels in the shipped configuration (full T = 1500): one ncu pass, no replay sets.
# ncu cannot launch cooperative cluster kernels: NABU_REC_NOCOOP=1 drops the attribute (co-residency then rests on the
# occupancy check, which holds on an otherwise idle GPU).
mkdir -p gpurun_out
export NABU_REC_NOCOOP=1
timeout -s KILL 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:"blstm_rec" -s 30 -c 10 --csv --log-file gpurun_out/traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit $? lines $(wc -l < gpurun_out/traffic.csv)"
python tools/ncu_traffic.py gpurun_out/traffic.csv dblstm_ctc > gpurun_out/r2_traffic.json; cat gpurun_out/r2_traffic.json | head -40
# launch list of one whole step (shares; cold-cache, serialised)
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv \
  --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1
echo "launch list exit $? lines $(wc -l < gpurun_out/launches_r2.csv)"
