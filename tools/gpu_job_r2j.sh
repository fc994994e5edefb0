#!/bin/bash
# round 2, call J: GPU suite (F64 greeks back in float64), fused VecNormalize after the fold / occupancy changes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log; tail -6 gpurun_out/r2j_pytest.log
timeout 300 python tools/bench_vecnorm_graph.py | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"hedge_step_kernel|vecnorm" --launch-skip 700 --launch-count 400 --csv --log-file gpurun_out/r2j_vn_launches.csv python tools/bench_vecnorm_graph.py > gpurun_out/r2j_vn_ncu.log 2>&1
python tools/ncu_launch_table.py gpurun_out/r2j_vn_launches.csv
