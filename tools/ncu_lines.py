#!/usr/bin/env python
"""Stall samples of one kernel aggregated by CUDA source line: joins ncu's SASS-level source page (instruction order) with
nvdisasm -g line info of the same object.  usage: ncu_lines.py report.ncu-rep object.o kernel-regex [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
name = [r[1] for r in rows if r and r[0] == "Kernel Name"][0] if any(r and r[0] == "Kernel Name" for r in rows) else None
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ci = {h: i for i, h in enumerate(hdr)}
body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) >= len(hdr):
        body.append(r)
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cubin = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
# find the section whose SASS matches the ncu rows: try every .text section of matching length
secs, cur = {}, None
for ln in dis:
    m = re.match(r"^\.text\.(\S+):", ln)
    if m:
        cur = m.group(1)
        secs[cur] = []
        loc = None
        continue
    if cur is None:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        loc = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        secs[cur].append((int(m.group(1), 16), loc, m.group(2)))
first_ops = [re.sub(r"\s+", " ", r[ci["Source"]].strip()).split(" ")[0] for r in body[:50]]
cands = [k for k, v in secs.items() if len(v) == len(body)]
if not cands:
    cands = sorted(secs, key=lambda k: abs(len(secs[k]) - len(body)))[:1]
sec = secs[cands[0]]
print(f"kernel section {cands[0][:80]}: {len(sec)} instructions (ncu: {len(body)})")
agg, cnt = collections.Counter(), collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
by_stall = collections.defaultdict(collections.Counter)
for (off, loc, sass), r in zip(sec, body):
    s = int(r[ci["# Samples"]] or 0)
    agg[loc] += s
    cnt[loc] += int(r[ci["Instructions Executed"]] or 0)
    for h in stall_cols:
        by_stall[loc][h[6:]] += int(r[ci[h]] or 0)
tot = sum(agg.values())
files = {}
def text(loc):
    if loc is None:
        return ""
    f = files.setdefault(loc[0], open(os.path.join(os.path.dirname(os.path.abspath(obj)), loc[0])).read().splitlines()
                         if os.path.exists(os.path.join(os.path.dirname(os.path.abspath(obj)), loc[0])) else [])
    return f[loc[1] - 1].strip()[:70] if 0 < loc[1] <= len(f) else ""
print(f"total samples {tot}")
for loc, s in agg.most_common(top):
    st = ", ".join(f"{k}:{v}" for k, v in by_stall[loc].most_common(3))
    print(f"{100.0 * s / max(tot, 1):5.1f}% {s:7d} inst {cnt[loc]:10d}  {loc[0] if loc else '?'}:{loc[1] if loc else 0:<4d} {text(loc):70s} [{st}]")
