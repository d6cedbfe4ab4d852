#!/bin/bash
# Round-2 last profiling pass: launch list of one cfg-3 step at the final build, ncu --set full of the persistent
# gemm_h2 kernel (x-projection NN, dX NT, weight gradients TN) and of the output layer's kernels (linear_skinny.cu).
# ncu cannot launch cooperative cluster kernels: NABU_REC_NOCOOP=1.
mkdir -p gpurun_out
export NABU_REC_NOCOOP=1
if [ -z "$SKIP_LAUNCHES" ]; then
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 372 -c 130 --csv \
  --log-file gpurun_out/launches_r2h_ctc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1
echo "launch list (cfg-3) exit $? lines $(wc -l < gpurun_out/launches_r2h_ctc.csv)"
fi
# per step 37 launches match (10 x-projections, output layer forward, dW + its reduction, 24 backward contractions): skip the
# three warm-up steps and layer 0's two small x-projections.  The report stays on the box (> 64 MiB with sources): the
# summary is made there.
timeout -s KILL 900 ncu --set full --clock-control none -k regex:"gemm_h2_kernel|linear_skinny" -s 113 -c 16 -o /tmp/r2h_full_gemm -f \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_full_gemm.log 2>&1
echo "gemm full exit $?"
python tools/ncu_summary.py full /tmp/r2h_full_gemm.ncu-rep > gpurun_out/r2h_ncu_gemm_linear.md 2> gpurun_out/ncu_summary.err
ncu -i /tmp/r2h_full_gemm.ncu-rep --page raw --csv > gpurun_out/r2h_ncu_gemm_linear_raw.csv 2>/dev/null
ls -la /tmp/*.ncu-rep gpurun_out/r2h_ncu_gemm_linear*; head -30 gpurun_out/r2h_ncu_gemm_linear.md | cut -c1-400
