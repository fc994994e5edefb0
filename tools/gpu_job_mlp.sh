CMD="python tools/bench_rollout.py --policies mlp_bf16 --sources gbm --reps 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_rollout_mlp_bf16_r1 $CMD > gpurun_out/ncu_mlp.log 2>&1; echo "ncu rc=$?"
python tools/ncu_lines.py gpurun_out/prof_rollout_mlp_bf16_r1.ncu-rep cantorrl_b200/csrc/rollout.o rollout_kernel 60 > gpurun_out/mlp_lines.txt 2>&1; tail -1 gpurun_out/mlp_lines.txt
python tools/ncu_summary.py gpurun_out/prof_rollout_mlp_bf16_r1.ncu-rep rollout_kernel 0 > gpurun_out/mlp_summary.txt 2>&1
