#!/bin/bash
# round 2, call Y: recurrent actor -- how many passes of a step read all 128 gate columns up front (0 / 1 / 2 = shipped / 3)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x -k "lstm or recurrent or shipped" 2>&1 | tail -2
for v in lstm_full0 lstm_full1 shipped lstm_full3; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm --reps 3 | grep -v "^{"
done
