#!/bin/bash
# parity tests + smoke only (a short GPU call)
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -40 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log
