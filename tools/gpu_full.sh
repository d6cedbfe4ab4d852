#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_ctc.json 2> gpurun_out/bench_ctc.err
timeout 600 python bench.py --workload las --no-cpu-baseline > gpurun_out/bench_las.json 2> gpurun_out/bench_las.err
if [ "${GPUS:-1}" -gt 1 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${GPUS} --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus ${GPUS} --steps 5 --warmup 3 > gpurun_out/bench_ctc_n${GPUS}.json 2> gpurun_out/bench_ctc_n${GPUS}.err
python tools/show_bench.py N${GPUS} < gpurun_out/bench_ctc_n${GPUS}.json; tail -3 gpurun_out/bench_ctc_n${GPUS}.err
fi
tail -4 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; python tools/show_bench.py ctc < gpurun_out/bench_ctc.json; python tools/show_bench.py las < gpurun_out/bench_las.json
