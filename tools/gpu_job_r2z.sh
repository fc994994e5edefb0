#!/bin/bash
# round 2, call Z: recurrent actor's issuer order: head layer first when ready (h), slot-0 refill after the product (r): variants h/r = 00 10 01 11
mkdir -p gpurun_out
for v in l00 l10 l01 shipped; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm --reps 3 | grep -v "^{"
done
CANTOR_HEDGE_LIB=build/variants/trace/libcantor_hedge.so timeout 120 python tools/lstm_trace.py > gpurun_out/r2z_trace.txt 2>&1
