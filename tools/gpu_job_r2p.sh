#!/bin/bash
# round 2, call P: new edge-case tests; ncu of the shipped persistent kernel; launch list of the bench command
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_step_modes_gpu.py tests/test_host_env_gpu.py -m gpu -q 2>&1 | tail -6
MIN="--e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --mlp-rollout-steps 0 --lstm-rollout-steps 0 --book-strikes 0 --rbergomi-paths 0 --l2free-envs 0 --no-forms"
timeout 600 ncu --set full --clock-control none --import-source on -f -k regex:hedge_step_many -s 2 -c 1 -o gpurun_out/r2p_prof_step_many python bench.py $MIN --steps 2 --warmup 1 > gpurun_out/r2p_ncu1.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/r2p_prof_step_many.ncu-rep "" 25 > gpurun_out/r2p_step_many_summary.txt 2>&1; head -12 gpurun_out/r2p_step_many_summary.txt | cut -c1-120
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/r2p_launches_bench.csv python bench.py $MIN --steps 2 --warmup 3 > gpurun_out/r2p_ncu2.log 2>&1; echo "ncu list rc=$?"
python tools/ncu_launch_table.py gpurun_out/r2p_launches_bench.csv
