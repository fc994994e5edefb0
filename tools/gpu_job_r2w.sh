#!/bin/bash
# round 2, call W: recurrent actor -- gate columns released as soon as they are in registers (+ pre-halved sigmoid rows): parity, A/B timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x -k "lstm or recurrent or shipped" 2>&1 | tail -3
for v in shipped lstm_late; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm,replay --reps 3 | grep -v "^{"
done
