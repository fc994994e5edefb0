python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
CMD="python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --rollout-steps 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -s 750 -c 300 --csv --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:hedge_step_kernel -s 800 -c 3 -f -o gpurun_out/prof_step_r1 $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/prof_rollout_r1 $CMD > gpurun_out/ncu3.log 2>&1
echo "ncu3 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:sim_paths_kernel -s 1 -c 1 -f -o gpurun_out/prof_sim_r1 $CMD > gpurun_out/ncu4.log 2>&1
echo "ncu4 rc=$?"
ls -la gpurun_out
