#!/bin/bash
# round 2, call 3V: DRAM bytes per launch of the per-step kernel at 2^23 envs with the alternating walk (60 consecutive launches inside a running sweep)
mkdir -p gpurun_out
timeout 600 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct \
  -k regex:hedge_step --launch-skip 300 --launch-count 60 --csv --log-file gpurun_out/r3v_step_2p23_dram.csv \
  env CANTOR_STEP_MANY_LAUNCHES=1 python tools/bench_step_l2free.py --sweeps 1 > gpurun_out/r3v_ncu.log 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/r3v_step_2p23_dram.csv")) if len(r) > 10]
h = rows[0]; im, iv, iu = h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = {}
for r in rows[1:]:
    v = float(r[iv].replace(",", ""))
    u = r[iu]
    if u == "Mbyte": v *= 1e6
    elif u == "Gbyte": v *= 1e9
    elif u == "Kbyte": v *= 1e3
    acc.setdefault(r[im], []).append(v)
for k, v in acc.items(): print(k, len(v), sum(v) / len(v))
PY
