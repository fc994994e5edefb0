set -x
timeout 300 python tools/bench_book.py > gpurun_out/book.log 2>&1; tail -2 gpurun_out/book.log
CMD1="python tools/bench_rollout.py --sources gbm --policies mlp_bf16 --reps 1"
timeout 200 $CMD1 > gpurun_out/plain_mlp.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 1 -c 1 -f -o gpurun_out/prof_rollout_mlp_bf16_r1 $CMD1 > gpurun_out/ncu5.log 2>&1
echo "ncu5 rc=$?"
CMD2="python tools/bench_book.py --reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:book_f32_kernel -s 1 -c 1 -f -o gpurun_out/prof_book_r1 $CMD2 > gpurun_out/ncu6.log 2>&1
echo "ncu6 rc=$?"
CMD3="python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --rollout-steps 1"
timeout 300 $CMD3 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/prof_rollout_r1 $CMD3 > gpurun_out/ncu3.log 2>&1
echo "ncu3 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sim_paths_kernel -s 1 -c 1 -f -o gpurun_out/prof_sim_r1 $CMD3 > gpurun_out/ncu4.log 2>&1
echo "ncu4 rc=$?"
