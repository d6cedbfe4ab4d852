mkdir -p gpurun_out
./tools/probe_tmem_a > gpurun_out/probe.txt 2>&1
NABU_BENCH_T=400 NABU_REC_TRACE=gpurun_out/trace timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/trace_bench.log 2>&1
python tools/trace_report.py gpurun_out/trace.fwd_tc.bin > gpurun_out/trace_fwd.txt 2>&1
python tools/trace_report.py gpurun_out/trace.bwd_tc.bin > gpurun_out/trace_bwd.txt 2>&1
NABU_BENCH_T=96 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"gemm|split|ctc|absmax|colsum|adam" -c 300 --csv --log-file gpurun_out/launches_norec_T96.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launch2.log 2>&1
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_kernels.py -x -q -k "test_blstm_fwd_bwd and 16-10-40-512" > gpurun_out/sanitizer.log 2>&1
cat gpurun_out/probe.txt; cat gpurun_out/trace_fwd.txt; tail -5 gpurun_out/sanitizer.log
