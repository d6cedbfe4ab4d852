#!/bin/bash
# Round-2 last profiling pass: launch list of one cfg-3 step at the final build, ncu --set full of the persistent
# gemm_h2 kernel (x-projection NN, dX NT, weight gradients TN) and of the output layer's kernels (linear_skinny.cu).
# ncu cannot launch cooperative cluster kernels: NABU_REC_NOCOOP=1.
mkdir -p gpurun_out
export NABU_REC_NOCOOP=1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 372 -c 130 --csv \
  --log-file gpurun_out/launches_r2h_ctc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1
echo "launch list (cfg-3) exit $? lines $(wc -l < gpurun_out/launches_r2h_ctc.csv)"
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_h2_kernel|linear_skinny" -s 108 -c 22 -o gpurun_out/r2h_full_gemm -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm full exit $?"
ls -la gpurun_out/*.ncu-rep
