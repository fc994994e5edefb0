#!/bin/bash
# round 2, call S: Monitor variants with lazily formed statistics: parity tests, then occupancy A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecnorm_gpu.py tests/test_step_modes_gpu.py tests/test_env_gpu.py -m gpu -q 2>&1 | tail -3
for v in shipped mon8 mon12; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"
  CANTOR_HEDGE_LIB=$lib VARIANTS=fp32_fast,fp32_monitor_stats,fp32_vecnormalize timeout 300 python tools/bench_step_variants.py 2>&1 | grep -v "^{"
done
