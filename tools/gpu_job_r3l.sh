#!/bin/bash
# round 2, call 3L: Monitor in the persistent step_many kernel -- running sums in registers, statistics formed lazily
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_modes_gpu.py tests/test_env_gpu.py tests/test_vecnorm_gpu.py -m gpu -q -x 2>&1 | tail -2
for v in shipped mon10; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_many_full.py --envs 1048576 --sweeps 6 --monitor 1 | tail -1
done
timeout 300 python tools/bench_many_full.py --envs 1048576 --sweeps 6 --monitor 1 --precision fp64 | tail -1
timeout 300 python tools/bench_many_full.py --envs 1048576 --sweeps 6 | tail -1
