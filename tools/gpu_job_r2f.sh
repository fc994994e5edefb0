#!/bin/bash
mkdir -p gpurun_out
MIN="--e2e-steps 0 --no-cpu-baseline --rollout-steps 0 --mlp-rollout-steps 0 --lstm-rollout-steps 0 --book-strikes 0 --rbergomi-paths 0 --l2free-envs 0 --no-forms"
one() { CANTOR_HEDGE_LIB=$1 timeout 300 python bench.py $MIN --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1]); print('$2 value=%.4e sweep_us=%.1f frac=%.3f' % (d['value'], d['roofline']['launch_us'], d['roofline']['frac']))"; }
for i in 1 2; do one cantorrl_b200/csrc/libcantor_hedge.so shipped_noef; one build/variants/many_ef/libcantor_hedge.so many_ef; done
echo "--- mlp rollout: warp-0 poller (shipped) vs all warps polling"
timeout 300 python tools/bench_rollout.py --policies mlp_bf16 --sources gbm,replay --reps 3 2>&1 | grep -v "^{" 
CANTOR_HEDGE_LIB=build/variants/mlp_allpoll/libcantor_hedge.so timeout 300 python tools/bench_rollout.py --policies mlp_bf16 --sources gbm,replay --reps 3 2>&1 | grep -v "^{"
timeout 300 python -m pytest tests/test_rollout_gpu.py -q -k "mlp" 2>&1 | tail -3
echo "--- vecnorm per-kernel times (ncu launch list)"
timeout 300 python tools/bench_vecnorm_graph.py > gpurun_out/r2f_vn_plain.log 2>&1; cat gpurun_out/r2f_vn_plain.log | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:"hedge_step_kernel|vecnorm" --launch-skip 200 --launch-count 120 --csv --log-file gpurun_out/r2f_vn_launches.csv python tools/bench_vecnorm_graph.py > gpurun_out/r2f_vn_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2f_vn_launches.csv")))
h = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[h]; ix = {k: i for i, k in enumerate(H)}
agg = collections.defaultdict(list)
for r in rows[h + 1:]:
    if len(r) >= len(H):
        agg[r[ix["Kernel Name"]][:70]].append(float(r[ix["Metric Value"]].replace(",", "")))
for k, v in agg.items():
    print(f"{k:72s} n={len(v):4d} mean={sum(v)/len(v)/1e3:8.2f} us  min={min(v)/1e3:8.2f}")
PY
