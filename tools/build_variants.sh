#!/bin/bash
# Build kernel-variant copies of libcantor_hedge.so under build/variants/ (travels to the GPU box, git-ignored).
# usage: tools/build_variants.sh "name1:-DFLAG=.. -DFLAG2=.." "name2:..."
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  d=build/variants/$name; mkdir -p $d
  for f in abi hedge_step formats path_sim reprice book_f32 host_env rollout vecnorm rbergomi; do
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
      $flags -Icantorrl_b200/csrc -c cantorrl_b200/csrc/$f.cu -o $d/$f.o -Xptxas -v 2> $d/$f.ptxas.log &
  done
  wait
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $d/libcantor_hedge.so $d/*.o
  echo "$name: $(grep -A1 'hedge_step_kernelILb0ELb0' $d/hedge_step.ptxas.log | grep -o 'Used [0-9]* registers' | head -1)"
done
