#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log; tail -8 gpurun_out/r2e_pytest.log
timeout 300 python tools/bench_modes.py --mode forms > gpurun_out/r2e_forms.log 2>&1; tail -1 gpurun_out/r2e_forms.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(json.dumps({k:d[k] for k in ('vecnormalize','step_record_info_fp32','step_monitor_fp32','step_fp64')}, indent=1))"
MIN="--e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --mlp-rollout-steps 0 --lstm-rollout-steps 0 --book-strikes 0 --rbergomi-paths 0 --l2free-envs 0 --no-forms"
timeout 300 python bench.py $MIN --steps 20 > gpurun_out/r2e_bench_min.log 2>&1; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2e_bench_min.log") if l.startswith("{")][-1])
print("shipped value=%.4e sweep_us=%.1f frac=%.3f per_step_launch_us=%.2f" % (d["value"], d["roofline"]["launch_us"], d["roofline"]["frac"], d["roofline"]["per_step_kernel"]["launch_us"]))
PY
