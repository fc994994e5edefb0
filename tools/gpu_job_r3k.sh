#!/bin/bash
# round 2, call 3K: recurrent actor with the 64 -> 2 output layer on the CUDA cores (one MMA round trip less per step) vs the three-MMA head
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in lstm_l3mma shipped; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm,replay --reps 3 | grep -v "^{"
done
CANTOR_HEDGE_LIB=build/variants/trace/libcantor_hedge.so timeout 120 python tools/lstm_trace.py > gpurun_out/r3k_trace.txt 2>&1
