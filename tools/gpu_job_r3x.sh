#!/bin/bash
# round 2, call 3X: MLP actor with K = 64 layers 2 / 3 (bias added in the epilogue: 9 instead of 11 MMAs per step) vs the K = 80 form; 6 / 7 CTAs per SM
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in mlp_k80 shipped mlp_k64_7; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_rollout.py --policies mlp_bf16 --sources gbm,heston,replay --reps 5 | grep -v "^{"
done
