#!/bin/bash
# Round-2 final profiling pass: DRAM traffic of the recurrences at the full T, launch lists of one cfg-3 step and one
# LAS step, ncu --set full of the decoder kernels.  ncu cannot launch cooperative cluster kernels: NABU_REC_NOCOOP=1.
mkdir -p gpurun_out
export NABU_REC_NOCOOP=1
timeout -s KILL 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
  -k regex:"blstm_rec" -s 30 -c 10 --csv --log-file gpurun_out/traffic.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit $? lines $(wc -l < gpurun_out/traffic.csv)"
python tools/ncu_traffic.py gpurun_out/traffic.csv dblstm_ctc > gpurun_out/r2f_traffic.json; head -30 gpurun_out/r2f_traffic.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 420 --csv \
  --log-file gpurun_out/launches_r2f_ctc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1
echo "launch list (cfg-3) exit $? lines $(wc -l < gpurun_out/launches_r2f_ctc.csv)"
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 11200 -c 3800 --csv \
  --log-file gpurun_out/launches_r2f_las.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload las > gpurun_out/ncu_launch_las.log 2>&1
echo "launch list (LAS) exit $? lines $(wc -l < gpurun_out/launches_r2f_las.csv)"
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"dec_attn_bwd_step|dec_attn_step|dec_lstm_step|dec_matmul_t" -s 1200 -c 8 -o gpurun_out/r2f_full_dec -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload las > gpurun_out/ncu_full_dec.log 2>&1
echo "dec full exit $?"
ls -la gpurun_out/*.ncu-rep
