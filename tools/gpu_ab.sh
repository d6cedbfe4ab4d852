#!/bin/bash
# same-box A/B: a reference build of the library (nabu_b200/csrc/build/lib_head.so -- build the commit to compare against in a
# git worktree and copy its libnabu_b200.so there; it must export every symbol lib.py binds) against the working tree's, then
# env variants of the latter
mkdir -p gpurun_out
cp nabu_b200/libnabu_b200.so /tmp/cur.so
run() {  # label, env...
  local label=$1; shift
  env "$@" timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ab_$label.json 2> gpurun_out/ab_$label.err
  echo "== $label: $(python tools/show_bench.py ctc < gpurun_out/ab_$label.json | sed -n 1p\;3p | cut -c1-220 | tr '\n' ' ')"
}
cp nabu_b200/csrc/build/lib_head.so nabu_b200/libnabu_b200.so; run head X=0
cp /tmp/cur.so nabu_b200/libnabu_b200.so; run cur X=0
IFS=';' read -ra VS <<< "${VARIANTS}"
i=0
for v in "${VS[@]}"; do run v$i $v; echo "   ($v)"; i=$((i+1)); done
cp nabu_b200/csrc/build/lib_head.so nabu_b200/libnabu_b200.so; run head2 X=0
cp /tmp/cur.so nabu_b200/libnabu_b200.so
