#!/bin/bash
# round 2, call 3P: gym-style step with 2 / 4 envs per thread (plain float32 fast path): parity, then timing at 2^16 .. 2^23 envs
mkdir -p gpurun_out
for e in 2 4; do
  echo "--- parity CANTOR_STEP_EPT=$e"; CANTOR_STEP_EPT=$e timeout 900 python -m pytest tests/test_env_gpu.py tests/test_step_modes_gpu.py tests/test_host_env_gpu.py -m gpu -q -x 2>&1 | tail -2
done
for n in 65536 262144 1048576 4194304 8388608; do
  for e in 1 2 4; do
    echo "--- envs $n ept $e"; CANTOR_STEP_EPT=$e timeout 300 python tools/bench_modes.py --mode replay --envs $n --sweeps 3 | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('%.2f us per step  frac %.3f' % (d['us_per_env_step_launch'], d['frac']))"
  done
done
