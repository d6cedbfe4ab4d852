#!/bin/bash
# trace + short bench per env variant (';'-separated VARIANTS), no tests
mkdir -p gpurun_out
IFS=';' read -ra VS <<< "${VARIANTS:-X=0}"
i=0
for v in "${VS[@]}"; do
  echo "== variant: $v"
  env $v NABU_BENCH_T=400 NABU_REC_TRACE=gpurun_out/trace$i timeout 300 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/trace_bench$i.log 2>&1
  python tools/trace_chains.py gpurun_out/trace$i.fwdc.bin 128 4 | sed -n 1p\;8p
  python tools/trace_chains.py gpurun_out/trace$i.bwd8c.bin 64 4 | sed -n 1p\;8p
  python tools/trace_blocks.py gpurun_out/trace$i.bwd8c.bin | head -1
  env $v timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/exp_ctc$i.json 2> gpurun_out/exp_ctc$i.err
  python tools/show_bench.py ctc < gpurun_out/exp_ctc$i.json | cut -c1-400 | sed -n 1p\;3p; tail -2 gpurun_out/exp_ctc$i.err
  i=$((i+1))
done
