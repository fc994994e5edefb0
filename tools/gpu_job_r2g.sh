#!/bin/bash
# round 2, call G: GPU suite with the float32 specials in the F64 greeks, fp64 step timing, per-kernel times of the fused VecNormalize
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log; tail -8 gpurun_out/r2g_pytest.log
echo "--- fp64 step with float32 specials"
timeout 300 python tools/bench_modes.py --mode replay --envs 8388608 --precision fp64 | tail -1
timeout 300 python tools/bench_modes.py --mode replay --envs 1048576 --precision fp64 | tail -1
echo "--- vecnorm per-kernel times (ncu launch list)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"hedge_step_kernel|vecnorm" --launch-skip 700 --launch-count 400 --csv --log-file gpurun_out/r2g_vn_launches.csv python tools/bench_vecnorm_graph.py > gpurun_out/r2g_vn_ncu.log 2>&1
python tools/ncu_launch_table.py gpurun_out/r2g_vn_launches.csv
