CMD="python tools/bench_rollout.py --policies delta_every_step --sources gbm --reps 1"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_rollout_r1 $CMD > gpurun_out/ncu_roll.log 2>&1; echo "ncu rc=$?"
python tools/ncu_lines.py gpurun_out/prof_rollout_r1.ncu-rep cantorrl_b200/csrc/rollout.o rollout_kernel 80 > gpurun_out/roll_lines.txt 2>&1; tail -1 gpurun_out/roll_lines.txt
python tools/ncu_summary.py gpurun_out/prof_rollout_r1.ncu-rep rollout_kernel 0 > gpurun_out/roll_summary.txt 2>&1
