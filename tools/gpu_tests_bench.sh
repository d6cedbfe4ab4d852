#!/bin/bash
# GPU suite + default bench line (+ optional extra command in $EXTRA)
mkdir -p gpurun_out
( timeout -s KILL ${PYTEST_TIMEOUT:-2400} python -m pytest tests -m gpu -q --durations=8 ${PYTEST_K:+-k "$PYTEST_K"} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|error" gpurun_out/pytest_gpu.log | tail -30
( timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS} > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ); echo "bench exit $?"
tail -5 gpurun_out/bench_default.err
python tools/show_bench.py < gpurun_out/bench_default.json 2>&1 | cut -c1-1200
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
