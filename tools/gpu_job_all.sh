timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench.log").read().strip().splitlines()[-1])
print("value", d["value"], "launch_us", d["roofline"]["launch_us"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
for k, v in d["extra"].items():
    print(k, {a: b for a, b in v.items() if a in ("ms", "ms_per_sweep", "env_steps_per_s", "path_steps_per_s", "reprices_per_s")})
PY
