#!/bin/bash
# compute-sanitizer over the kernels written in the last session of round 2: persistent gemm_h2 (BK = 32, SWIZZLE_64B, epilogue
# under the next tile's main loop, absmax output), linear_skinny.cu, the max |dx| hand-over of the BLSTM backward
mkdir -p gpurun_out
export NABU_REC_NOCOOP=1 NABU_QUIET=1
K='gemm or linear or planes'
for tool in memcheck racecheck synccheck; do
  ( timeout -s KILL ${SAN_TIMEOUT:-400} compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 600 \
      python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$K" ) > gpurun_out/r2h_sanitizer_$tool.log 2>&1
  echo "$tool exit $?"; tail -4 gpurun_out/r2h_sanitizer_$tool.log
done
