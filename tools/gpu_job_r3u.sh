#!/bin/bash
# round 2, call 3U: per-step replay kernel walking its env tiles in alternating directions (L2 reuse across launches), sizes around the L2 capacity
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_env_gpu.py tests/test_step_modes_gpu.py tests/test_vecnorm_gpu.py tests/test_host_env_gpu.py -m gpu -q -x 2>&1 | tail -2
for n in 1048576 2097152 4194304 8388608; do
  for v in noalt shipped; do
    lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
    echo "--- envs $n $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_modes.py --mode replay --envs $n --sweeps 3 | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('%.2f us per step  frac %.3f' % (d['us_per_env_step_launch'], d['frac']))"
  done
done
for v in noalt shipped; do lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so; echo "--- vecnorm $v"; CANTOR_HEDGE_LIB=$lib python tools/bench_vecnorm_graph.py 2>&1 | tail -1; done
