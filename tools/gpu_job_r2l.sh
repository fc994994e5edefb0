#!/bin/bash
# round 2, call L: two-tile tensor-core MLP rollout: parity tests, then A/B against the one-tile form
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; tail -15 gpurun_out/r2l_pytest.log
echo "--- two-tile (shipped)"
timeout 300 python tools/bench_rollout.py --policies mlp_bf16 --sources gbm,heston,replay --reps 3 2>&1 | grep -v "^{"
timeout 300 python tools/bench_rollout.py --policies mlp_bf16 --sources gbm --reps 2 --envs 524288 --steps 1000 2>&1 | grep -v "^{"
echo "--- one-tile (round 1)"
CANTOR_MLP_ONE_TILE=1 timeout 300 python tools/bench_rollout.py --policies mlp_bf16 --sources gbm,heston,replay --reps 3 2>&1 | grep -v "^{"
CANTOR_MLP_ONE_TILE=1 timeout 300 python tools/bench_rollout.py --policies mlp_bf16 --sources gbm --reps 2 --envs 524288 --steps 1000 2>&1 | grep -v "^{"
