timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.log; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --precision fp64 --steps 10 --e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --mlp-rollout-steps 0 --book-strikes 0 --rbergomi-paths 0 > gpurun_out/bench_fp64.log 2>&1
CMD3="python bench.py --steps 1 --warmup 3 --e2e-steps 0 --no-cpu-baseline --rollout-steps 1 --mlp-rollout-steps 0 --book-strikes 0 --rbergomi-paths 0"
timeout 300 $CMD3 > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 750 -c 300 --csv --log-file gpurun_out/launches_r1.csv $CMD3 > gpurun_out/ncu1.log 2>&1
echo "ncu1 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hedge_step_kernel -s 800 -c 3 -f -o gpurun_out/prof_step_r1 $CMD3 > gpurun_out/ncu2.log 2>&1
echo "ncu2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/prof_rollout_r1 $CMD3 > gpurun_out/ncu3.log 2>&1
echo "ncu3 rc=$?"
CMD2="python tools/bench_book.py --reps 1"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:book_f32_kernel -s 1 -c 1 -f -o gpurun_out/prof_book_r1 $CMD2 > gpurun_out/ncu6.log 2>&1
echo "ncu6 rc=$?"
timeout 300 python tools/bench_step_variants.py > gpurun_out/step_variants.log 2>&1; tail -1 gpurun_out/step_variants.log | cut -c1-300
timeout 300 python tools/bench_rollout.py > gpurun_out/rollout_variants.log 2>&1; tail -1 gpurun_out/rollout_variants.log | cut -c1-300
