#!/bin/bash
# round 2, call 3D: persistent step_many -- per-warp observation stores (no CTA barrier in a step) vs per-CTA stores; occupancy / prefetch on top
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_step_modes_gpu.py tests/test_env_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in cta_stores shipped ws12 ws_pf2; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_many_full.py --envs 1048576 --sweeps 10 | tail -1
done
echo "--- shipped, monitor / fp64"
timeout 300 python tools/bench_many_full.py --envs 1048576 --sweeps 5 --monitor 1 | tail -1
timeout 300 python tools/bench_many_full.py --envs 1048576 --sweeps 5 --precision fp64 | tail -1
