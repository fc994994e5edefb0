#!/bin/bash
# round 2, call 3E: VecNormalize fused form with the fold inside the apply kernel (one launch less per env-step) vs the separate fold launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vecnorm_gpu.py -m gpu -q -x 2>&1 | tail -2
CANTOR_VECNORM_SEPARATE_FOLD=1 timeout 900 python -m pytest tests/test_vecnorm_gpu.py -m gpu -q -x 2>&1 | tail -2
echo "--- fold inside apply"; timeout 300 python tools/bench_vecnorm_graph.py | tail -1
echo "--- separate fold launch"; CANTOR_VECNORM_SEPARATE_FOLD=1 timeout 300 python tools/bench_vecnorm_graph.py | tail -1
