#!/bin/bash
# round 2, call 3Y: DRAM bytes of pack / unpack book, path-fastest 2-D tile order (before) vs time-fastest 1-D order (shipped)
mkdir -p gpurun_out
for v in fmt_old shipped; do
  lib=build/variants/$v/libcantor_hedge.so; [ $v = shipped ] && lib=cantorrl_b200/csrc/libcantor_hedge.so
  CANTOR_HEDGE_LIB=$lib timeout 300 ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:"pack_book" --launch-skip 2 --launch-count 4 --csv --log-file gpurun_out/r3y_$v.csv python tools/fmt_probe.py > gpurun_out/r3y_$v.log 2>&1
  echo "--- $v"; python - <<PY
import csv
rows = [r for r in csv.reader(open("gpurun_out/r3y_$v.csv")) if len(r) > 10]
h = rows[0]; ik, im, iv, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = {}
for r in rows[1:]:
    v = float(r[iv].replace(",", "")); u = r[iu]
    v *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3}.get(u, 1.0)
    acc.setdefault((r[ik][:24], r[im]), []).append(v)
for k, v in sorted(acc.items()): print("  %-26s %-24s %12.1f" % (k[0], k[1], (sum(v) / len(v)) / (1e6 if "bytes" in k[1] else 1e3)), "MB" if "bytes" in k[1] else "us")
PY
done
