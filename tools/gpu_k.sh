#!/bin/bash
# run a subset of the GPU tests: PYTEST_K='expr' tools/gpu_k.sh
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -q -k "$PYTEST_K" ) > gpurun_out/pytest_gpu_k.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_k.log
tail -30 gpurun_out/pytest_gpu_k.log
