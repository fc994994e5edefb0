#!/bin/bash
# round 2, call Q: smoke with the three step forms; HBM context: pure-write, pure-read and copy bandwidth as torch measures them
mkdir -p gpurun_out
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -2
python - <<'PY'
import torch
n = 1 << 32                                   # 8 GiB of bf16
a = torch.empty(n, dtype=torch.bfloat16, device="cuda")
b = torch.empty(n, dtype=torch.bfloat16, device="cuda")
def best(fn, nbytes, reps=10):
    t = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t.append(e0.elapsed_time(e1))
    return nbytes / min(t) / 1e6
print("copy  (read + write bytes) %.1f GB/s" % best(lambda: b.copy_(a), 2 * n * 2))
print("fill  (write only)         %.1f GB/s" % best(lambda: a.fill_(1.0), n * 2))
print("sum   (read only)          %.1f GB/s" % best(lambda: a.view(torch.int32).sum(), n * 2))
c = torch.empty(n // 2, dtype=torch.bfloat16, device="cuda")
print("2 reads + 1... add_ (read 2, write 1: 33%% writes) %.1f GB/s" % best(lambda: torch.add(a[: n // 2], b[: n // 2], out=c), 3 * (n // 2) * 2))
PY
