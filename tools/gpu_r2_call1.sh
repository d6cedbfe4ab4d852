#!/bin/bash
# round 2, call 1: the whole GPU suite with the new BASELINE-size parity tests, the reworked bench line, H=128 probe
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
( timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
( timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err ); echo "bench exit $?"
tail -5 gpurun_out/bench_default.err
python tools/show_bench.py < gpurun_out/bench_default.json 2>&1 | head -60
( timeout 300 python tools/h128_probe.py ) > gpurun_out/h128_probe.txt 2>&1
cat gpurun_out/h128_probe.txt
