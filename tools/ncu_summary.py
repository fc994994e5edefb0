#!/usr/bin/env python
"""Summarise an .ncu-rep (run where ncu is installed): key metrics per launch + SASS opcode mix + stall reasons."""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
pat = sys.argv[2] if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "lts__t_sector_hit_rate.pct",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "launch__occupancy_limit_blocks", "sm__maximum_warps_per_active_cycle_pct",
        "launch__grid_size", "launch__block_size"]
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            vals = [r[i][:60] for r in data]
            print(f"{w} [{units[i]}]: {vals}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--kernel-name", f"regex:{pat}"] if pat else []),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) >= len(hdr):
        body.append(r)
tot = sum(int(r[ci["Instructions Executed"]] or 0) for r in body)
op, samp = collections.Counter(), collections.Counter()
for r in body:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
    o = m.group(2).split(".")[0] if m else "?"
    op[o] += int(r[ci["Instructions Executed"]] or 0)
    samp[o] += int(r[ci["# Samples"]] or 0)
print(f"SASS instructions: {len(body)} static, {tot} warp-instructions executed")
print("opcode mix (warp-instr, % of total, stall samples):")
for o, c in op.most_common(24):
    print(f"  {o:10s} {c:12d} {100.0 * c / max(tot, 1):5.1f}%  samples {samp[o]}")
st = {k: sum(int(r[ci[k]] or 0) for r in body) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}
print("stall samples:", sorted(st.items(), key=lambda kv: -kv[1])[:8])

# hottest SASS instructions by stall samples (with the CUDA source line when the build has -lineinfo)
top = sorted(body, key=lambda r: -int(r[ci["# Samples"]] or 0))[:int(sys.argv[3]) if len(sys.argv) > 3 else 0]
for r in top:
    print(f"  {int(r[ci['# Samples']] or 0):7d}  {r[ci['Address']][-6:]}  {r[ci['Source']].strip()[:90]}")
