#!/bin/bash
# bench variants: one line each
mkdir -p gpurun_out
run() { tag=$1; shift; ( env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-extras --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/var_$tag.json 2> gpurun_out/var_$tag.err ); python tools/show_bench.py "$tag" < gpurun_out/var_$tag.json 2>&1 | head -4 | cut -c1-900; tail -2 gpurun_out/var_$tag.err; }
for spec in "$@"; do
  tag=${spec%%:*}; envs=${spec#*:}
  run $tag $(echo $envs | tr ',' ' ')
done
