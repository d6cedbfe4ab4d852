#!/bin/bash
# the whole GPU suite (no -x: every failure is listed), slowest tests, parity tables
mkdir -p gpurun_out
( timeout -s KILL ${PYTEST_TIMEOUT:-2400} python -m pytest tests -m gpu -q --durations=25 ${PYTEST_K:+-k "$PYTEST_K"} ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
