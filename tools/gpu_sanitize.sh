#!/bin/bash
# compute-sanitizer over the kernel-level GPU tests at their small shapes (cluster / tcgen05 kernels included)
mkdir -p gpurun_out
export NABU_REC_NOCOOP=1 NABU_QUIET=1
K='test_blstm_fwd_bwd or test_blstm_padded or planes or test_ctc or test_linear or test_clip or test_speller_fwd_bwd or test_las_beam or test_ctc_beam or attention_entry'
( timeout -s KILL ${SAN_TIMEOUT:-1500} compute-sanitizer --tool memcheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_kernels.py tests/test_gpu_speller.py tests/test_gpu_decoders.py tests/test_gpu_attention_api.py -m gpu -q -x -k "$K" ) > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck exit $?"; tail -5 gpurun_out/sanitizer_memcheck.log
( timeout -s KILL ${SAN_TIMEOUT:-1500} compute-sanitizer --tool initcheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "test_blstm_fwd_bwd or planes" ) > gpurun_out/sanitizer_initcheck.log 2>&1
echo "initcheck exit $?"; tail -5 gpurun_out/sanitizer_initcheck.log
# racecheck (shared-memory hazards) and synccheck (barrier misuse) over the recurrence kernels at small shapes
( timeout -s KILL ${SAN_TIMEOUT:-1500} compute-sanitizer --tool racecheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "test_blstm_fwd_bwd or planes or h1024" ) > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck exit $?"; tail -6 gpurun_out/sanitizer_racecheck.log
( timeout -s KILL ${SAN_TIMEOUT:-1500} compute-sanitizer --tool synccheck --error-exitcode 9 --launch-timeout 600 \
    python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "test_blstm_fwd_bwd or planes or h1024" ) > gpurun_out/sanitizer_synccheck.log 2>&1
echo "synccheck exit $?"; tail -6 gpurun_out/sanitizer_synccheck.log
