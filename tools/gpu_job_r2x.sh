#!/bin/bash
# round 2, call X: recurrent actor with the groups half a step apart (8 half-passes, 2 gate halves, release at load): parity, timing, timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm,replay --reps 3 | grep -v "^{"
timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm --reps 2 --store --envs 262144 | grep -v "^{"
CANTOR_HEDGE_LIB=build/variants/trace/libcantor_hedge.so timeout 120 python tools/lstm_trace.py > gpurun_out/r2x_trace.txt 2>&1; tail -3 gpurun_out/r2x_trace.txt
