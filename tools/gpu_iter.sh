#!/bin/bash
# kernel iteration pass: BLSTM parity tests (PYTEST_K overrides), cfg-3 bench without extras, optional LAS line (LAS=1)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q -k "${PYTEST_K:-blstm or dblstm or cfg3}" ) > gpurun_out/pytest_gpu_k.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_k.log
tail -4 gpurun_out/pytest_gpu_k.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/iter_ctc.json 2> gpurun_out/iter_ctc.err
python tools/show_bench.py ctc < gpurun_out/iter_ctc.json | cut -c1-700; tail -3 gpurun_out/iter_ctc.err
if [ -n "$LAS" ]; then
  timeout 300 python bench.py --workload las --no-cpu-baseline --no-extras > gpurun_out/iter_las.json 2> gpurun_out/iter_las.err
  python tools/show_bench.py las < gpurun_out/iter_las.json | cut -c1-900; tail -3 gpurun_out/iter_las.err
fi
