#!/bin/bash
# ncu --set full of the two recurrence kernels (T = 96 so that the ~40 replays stay short) + the LAS attention kernels.
# ncu cannot launch cooperative cluster kernels: NABU_REC_NOCOOP=1 (see tools/gpu_ncu_traffic.sh).
mkdir -p gpurun_out
export NABU_REC_NOCOOP=1 NABU_OVERLAP=0 NABU_BENCH_T=96
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"blstm_rec_fwd" -s 7 -c 1 -o gpurun_out/r2_full_rec_fwd -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_rec_fwd.log 2>&1
echo "fwd exit $?"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"blstm_rec_bwd" -s 7 -c 1 -o gpurun_out/r2_full_rec_bwd -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_rec_bwd.log 2>&1
echo "bwd exit $?"
unset NABU_BENCH_T
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"dec_attn" -s 400 -c 2 -o gpurun_out/r2_full_dec -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --workload las > gpurun_out/ncu_full_dec.log 2>&1
echo "dec exit $?"
ls -la gpurun_out/*.ncu-rep
