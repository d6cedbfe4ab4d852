#!/bin/bash
# LAS iteration pass: decoder / attention / golden parity tests (PYTEST_K overrides), LAS bench line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "${PYTEST_K:-speller or las or cfg2 or cfg4 or tf18 or golden or attention or decoder}" ) > gpurun_out/pytest_gpu_las.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_las.log
tail -6 gpurun_out/pytest_gpu_las.log
timeout 300 python bench.py --workload las --no-cpu-baseline --no-extras > gpurun_out/iter_las.json 2> gpurun_out/iter_las.err
python tools/show_bench.py las < gpurun_out/iter_las.json | cut -c1-1200; tail -3 gpurun_out/iter_las.err
if [ -n "$EXTRA" ]; then bash -c "$EXTRA"; fi
