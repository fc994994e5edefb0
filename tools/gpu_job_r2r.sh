#!/bin/bash
# round 2, call R: MLP actor variants: last layer on the CUDA cores, FFMA normalisation
mkdir -p gpurun_out
for v in shipped mlp_head_cc mlp_fma_norm mlp_both; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"
  CANTOR_HEDGE_LIB=$lib timeout 300 python -m pytest tests/test_rollout_gpu.py -m gpu -q -k "mlp" 2>&1 | tail -2
  CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_rollout.py --policies mlp_bf16 --sources gbm,replay --reps 3 2>&1 | grep -v "^{"
done
