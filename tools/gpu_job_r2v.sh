#!/bin/bash
# round 2, call V: segment-major work units of the persistent step_many kernel: parity (bit-identical to per-step launches), then timing
mkdir -p gpurun_out
for k in 1 4 7; do
  echo "--- parity, CANTOR_MANY_SEGMENTS=$k"
  CANTOR_MANY_SEGMENTS=$k timeout 600 python -m pytest tests/test_step_modes_gpu.py -m gpu -q -x 2>&1 | tail -3
done
for k in 1 2 3 4 6 9 12 21; do
  echo "--- CANTOR_MANY_SEGMENTS=$k"; CANTOR_MANY_SEGMENTS=$k timeout 300 python tools/bench_many_full.py --envs 1048576 --sweeps 10 | tail -1
done
echo "--- whole waves, one segment"; timeout 300 python tools/bench_many_full.py --envs 947200,1136640 --sweeps 10
