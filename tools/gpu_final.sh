#!/bin/bash
# round-end pass on one B200: full parity suite, smoke, default bench line (with cpu_baseline), LAS bench line
mkdir -p gpurun_out
( time timeout 300 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 300 python bench.py > gpurun_out/bench_ctc.json 2> gpurun_out/bench_ctc.err
timeout 200 python bench.py --workload las --no-cpu-baseline > gpurun_out/bench_las.json 2> gpurun_out/bench_las.err
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; python tools/show_bench.py ctc < gpurun_out/bench_ctc.json; python tools/show_bench.py las < gpurun_out/bench_las.json
