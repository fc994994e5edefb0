#!/bin/bash
# round 2, call T: occupancy of the moments-producing step kernel; full-shard agreement test
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_modes_gpu.py -m gpu -q -k "full_shard" 2>&1 | tail -3
for v in shipped vn14 vn16; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  echo "--- $v"; CANTOR_HEDGE_LIB=$lib timeout 300 python tools/bench_vecnorm_graph.py | tail -1
done
CANTOR_HEDGE_LIB=build/variants/vn16/libcantor_hedge.so timeout 300 python -m pytest tests/test_vecnorm_gpu.py -m gpu -q 2>&1 | tail -2
