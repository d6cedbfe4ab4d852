#!/bin/bash
# quick GPU pass: BLSTM + model parity tests, cfg-3 bench, phase traces
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ${PYTEST_K:+-k "$PYTEST_K"} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ctc.json 2> gpurun_out/bench_ctc.err
NABU_BENCH_T=400 NABU_REC_TRACE=gpurun_out/trace timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/trace_bench.log 2>&1
python tools/trace_report.py gpurun_out/trace.fwd_tc.bin > gpurun_out/trace_fwd.txt 2>&1
python tools/trace_report.py gpurun_out/trace.bwd8.bin > gpurun_out/trace_bwd.txt 2>&1
tail -4 gpurun_out/pytest_gpu.log; python tools/show_bench.py < gpurun_out/bench_ctc.json; cat gpurun_out/bench_ctc.err | tail -5
head -12 gpurun_out/trace_fwd.txt; head -12 gpurun_out/trace_bwd.txt
