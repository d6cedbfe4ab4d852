#!/bin/bash
# round-end pass on one B200: full parity suite, smoke, default bench line (with cpu_baseline, las and decode objects)
mkdir -p gpurun_out
( time timeout 400 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench exit $?"
tail -6 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; python tools/show_bench.py < gpurun_out/bench_default.json | cut -c1-600
