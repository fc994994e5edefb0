CMD="python tools/bench_rollout.py --envs 37888 --policies lstm_bf16 --sources gbm --reps 1"
timeout 100 $CMD | tail -1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_lstm_r1 $CMD > gpurun_out/ncu_lstm.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py gpurun_out/prof_lstm_r1.ncu-rep rollout_kernel 0 > gpurun_out/lstm_summary.txt 2>&1
python tools/ncu_lines.py gpurun_out/prof_lstm_r1.ncu-rep cantorrl_b200/csrc/rollout.o rollout_kernel 70 > gpurun_out/lstm_lines.txt 2>&1; tail -3 gpurun_out/lstm_lines.txt
