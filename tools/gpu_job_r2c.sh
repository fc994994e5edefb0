#!/bin/bash
# round 2, call C: whole GPU suite again, persistent-kernel A/B (plain stores vs TMA + fence), ncu captures of the new kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log; tail -12 gpurun_out/r2c_pytest.log
MIN="--e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --mlp-rollout-steps 0 --lstm-rollout-steps 0 --book-strikes 0 --rbergomi-paths 0 --l2free-envs 0 --no-forms"
timeout 300 python bench.py $MIN --steps 20 > gpurun_out/r2c_bench_min.log 2>&1; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2c_bench_min.log") if l.startswith("{")][-1])
print("shipped value=%.4e sweep_us=%.1f frac=%.3f per_step_launch_us=%.2f" % (d["value"], d["roofline"]["launch_us"], d["roofline"]["frac"], d["roofline"]["per_step_kernel"]["launch_us"]))
PY
timeout 600 python tools/sweep_variants.py --steps 20 > gpurun_out/r2c_variants.log 2>&1; cat gpurun_out/r2c_variants.log
timeout 300 python tools/bench_modes.py --mode many --envs 8388608 > gpurun_out/r2c_many_2p23.log 2>&1; tail -1 gpurun_out/r2c_many_2p23.log
# ---- ncu: full captures (one launch each) ----
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 600 $NCU -k regex:hedge_step_many -s 2 -c 1 -o gpurun_out/r2c_prof_step_many python bench.py $MIN --steps 2 --warmup 1 > gpurun_out/r2c_ncu1.log 2>&1; echo "ncu many rc=$?"
timeout 600 $NCU -k regex:hedge_step_sim -s 300 -c 1 -o gpurun_out/r2c_prof_step_sim python tools/bench_modes.py --mode sim --envs 8388608 --sweeps 1 > gpurun_out/r2c_ncu2.log 2>&1; echo "ncu sim rc=$?"
timeout 600 $NCU -k regex:hedge_step_kernel -s 300 -c 1 -o gpurun_out/r2c_prof_step_f64 python tools/bench_modes.py --mode replay --envs 1048576 --precision fp64 --sweeps 1 > gpurun_out/r2c_ncu3.log 2>&1; echo "ncu f64 rc=$?"
timeout 600 $NCU -k regex:schema_b_book -s 1 -c 1 -o gpurun_out/r2c_prof_schema_b python tools/bench_modes.py --mode forms > gpurun_out/r2c_ncu4.log 2>&1; echo "ncu schema rc=$?"
# ---- ncu: DRAM bytes of the persistent kernel with caches left alone (traffic line of the bench) ----
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
  -k regex:hedge_step_many -s 2 -c 3 --csv --log-file gpurun_out/r2c_step_many_dram.csv python bench.py $MIN --steps 4 --warmup 1 > gpurun_out/r2c_ncu5.log 2>&1
for r in step_many step_sim step_f64 schema_b; do python tools/ncu_summary.py gpurun_out/r2c_prof_$r.ncu-rep > gpurun_out/r2c_${r}_summary.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep | tail -5
