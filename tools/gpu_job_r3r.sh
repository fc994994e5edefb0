#!/bin/bash
# round 2, call 3R: recurrent actor in a 12-warp CTA with setmaxnreg (env warps 224 registers, issuer warpgroup 56): whole-pass gate reads 0 / 2 / 3 passes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in full0 full2 shipped; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm,replay --reps 3 | grep -v "^{"
done
timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm --reps 2 --store --envs 262144 | grep -v "^{"
CANTOR_HEDGE_LIB=build/variants/trace/libcantor_hedge.so timeout 120 python tools/lstm_trace.py > gpurun_out/r3r_trace.txt 2>&1
