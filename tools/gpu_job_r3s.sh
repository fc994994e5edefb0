#!/bin/bash
# round 2, call 3S: recurrent actor with one issuer warp per group (shared weight ring handed over on an mbarrier)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x 2>&1 | tail -3
timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm,replay --reps 3 | grep -v "^{"
timeout 300 python tools/bench_rollout.py --policies lstm_bf16 --sources gbm --reps 2 --store --envs 262144 | grep -v "^{"
CANTOR_HEDGE_LIB=build/variants/trace/libcantor_hedge.so timeout 120 python tools/lstm_trace.py > gpurun_out/r3s_trace.txt 2>&1; tail -2 gpurun_out/r3s_trace.txt
