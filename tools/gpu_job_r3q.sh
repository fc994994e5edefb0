#!/bin/bash
# round 2, call 3A (second session): what the driver runs at round end, on one GPU: GPU suite, smoke, both bench arms with the driver's flags
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r3q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3q_pytest.log; tail -8 gpurun_out/r3q_pytest.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r3q_smoke.log 2>&1; tail -1 gpurun_out/r3q_smoke.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r3q_bench_ref.log 2> gpurun_out/r3q_bench_ref.err; echo "ref rc=$?"; tail -3 gpurun_out/r3q_bench_ref.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/r3q_bench.log 2> gpurun_out/r3q_bench.err; echo "bench rc=$?"; tail -4 gpurun_out/r3q_bench.err
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r3q_bench.log") if l.startswith("{")][-1])
r = json.loads([l for l in open("gpurun_out/r3q_bench_ref.log") if l.startswith("{")][-1])
print("value %.4e frac %.3f e2e %.4e  ref %.4e  e2e_ratio %.0f  same_config %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], r["value"], d["e2e"]["value"] / r["value"], d["config"] == r["config"]))
print("traffic/env-step", d["roofline"]["traffic_bytes_per_env_step"], "l2free", d["roofline"]["per_step_kernel_l2_free"])
print("clocks", d["clocks"])
PY
