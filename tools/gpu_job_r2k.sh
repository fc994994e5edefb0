#!/bin/bash
# round 2, call K: fused VecNormalize with the returns requested at kernel start
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vecnorm_gpu.py -m gpu -q 2>&1 | tail -3
timeout 300 python tools/bench_vecnorm_graph.py | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"hedge_step_kernel|vecnorm" --launch-skip 700 --launch-count 400 --csv --log-file gpurun_out/r2k_vn_launches.csv python tools/bench_vecnorm_graph.py > gpurun_out/r2k_vn_ncu.log 2>&1
python tools/ncu_launch_table.py gpurun_out/r2k_vn_launches.csv
