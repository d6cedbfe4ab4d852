#!/bin/bash
# phase traces of the chain recurrences at the current build (cfg-3 shapes, T = 400), no tests
mkdir -p gpurun_out
NABU_BENCH_T=400 NABU_REC_TRACE=gpurun_out/trace timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/trace_bench.log 2>&1


python tools/trace_chains.py gpurun_out/trace.fwdc.bin 128 4; python tools/trace_chains.py gpurun_out/trace.bwd8c.bin 64 4
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
