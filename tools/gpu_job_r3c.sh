#!/bin/bash
# round 2, call 3C: persistent step_many with its CTAs in clusters that stay within one step of each other (split cluster barrier)
mkdir -p gpurun_out
for v in cl8; do
  echo "--- parity $v"; CANTOR_HEDGE_LIB=build/variants/$v/libcantor_hedge.so timeout 600 python -m pytest tests/test_step_modes_gpu.py -m gpu -q -x 2>&1 | tail -2
done
for v in shipped cl2 cl4 cl8 cl8_8; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_many_full.py --envs 1048576 --sweeps 10 | tail -1
done
