#!/bin/bash
# compute-sanitizer over the decoder kernels (cluster attention backward, bulk-copy staged matmuls, dependent-launch chain)
mkdir -p gpurun_out
export NABU_QUIET=1
K='test_speller_fwd_bwd or test_speller_dropout or test_las_beam or attention'
for tool in memcheck racecheck synccheck; do
  ( timeout -s KILL ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --error-exitcode 9 --launch-timeout 600 \
      python -m pytest tests/test_gpu_speller.py tests/test_gpu_attention_api.py -m gpu -q -x -k "$K" ) > gpurun_out/sanitizer_decoder_$tool.log 2>&1
  echo "$tool exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/sanitizer_decoder_$tool.log | tail -3
done
