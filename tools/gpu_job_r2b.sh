#!/bin/bash
# round 2, call B: whole GPU test suite, smoke, the new bench line (both arms), persistent-kernel variants
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log; tail -15 gpurun_out/r2b_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1; tail -1 gpurun_out/r2b_smoke.log
timeout 900 python bench.py > gpurun_out/r2b_bench.log 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2b_bench.log; tail -5 gpurun_out/r2b_bench.err
timeout 400 python bench.py --impl reference --steps 10 --warmup 2 > gpurun_out/r2b_bench_ref.log 2> gpurun_out/r2b_bench_ref.err; echo "ref rc=$?"; tail -c 700 gpurun_out/r2b_bench_ref.log
timeout 600 python tools/sweep_variants.py --steps 20 > gpurun_out/r2b_variants.log 2>&1; cat gpurun_out/r2b_variants.log
