#!/bin/bash
# round 2, call 3W: on-the-fly step with the alternating walk (CANTOR_STEP_WALK_BACKWARD on odd steps): parity, timing at 2^23 (GBM) / 2^24 (Heston)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_modes_gpu.py tests/test_vecnorm_gpu.py tests/test_env_gpu.py -m gpu -q -x 2>&1 | tail -2
timeout 300 python tools/bench_modes.py --mode sim --envs 8388608 --sweeps 2 | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('gbm 2^23: %.2f us per step  frac %.3f' % (d['us_per_env_step_launch'], d['frac']))"
timeout 300 python tools/bench_modes.py --mode sim --envs 16777216 --model heston --sweeps 2 | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.read()); print('heston 2^24: %.2f us per step  frac %.3f' % (d['us_per_env_step_launch'], d['frac']))"
