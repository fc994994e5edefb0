#!/usr/bin/env python
"""HBM bandwidth of this GPU as a function of the read / write mix, by plain torch streaming kernels over multi-GiB buffers
(context for roofline fractions quoted against the 50 % / 50 % copy peak)."""
import torch

dev = torch.device("cuda", 0)
G = 1 << 30


def best(fn, nbytes, reps=8):
    fn(); torch.cuda.synchronize()
    out = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        out = min(out, e0.elapsed_time(e1))
    return nbytes / out / 1e6


f32 = torch.empty(2 * G // 4, dtype=torch.float32, device=dev).normal_()      # 2 GiB
f64 = torch.empty(2 * G // 4, dtype=torch.float64, device=dev)                # 4 GiB
f32b = torch.empty_like(f32)
h16 = torch.empty(2 * G // 4, dtype=torch.float16, device=dev)                # 1 GiB
print(f"  0 % writes  sum(float32)                 {best(lambda: f32.sum(), 2 * G):8.1f} GB/s")
print(f" 33 % writes  float64 -> float32           {best(lambda: f32b.copy_(f64), 6 * G):8.1f} GB/s")
print(f" 33 % writes  add(a, b, out=c) float32     {best(lambda: torch.add(f32, f32b, out=f64.view(torch.float32)[: f32.numel()]), 6 * G):8.1f} GB/s")
print(f" 50 % writes  copy float32                 {best(lambda: f32b.copy_(f32), 4 * G):8.1f} GB/s")
print(f" 67 % writes  float32 -> float64           {best(lambda: f64.copy_(f32), 6 * G):8.1f} GB/s")
print(f" 80 % writes  float16 -> float64           {best(lambda: f64.copy_(h16), 5 * G):8.1f} GB/s")
print(f"100 % writes  fill_(1.0) float32           {best(lambda: f64.view(torch.float32).fill_(1.0), 4 * G):8.1f} GB/s")
print(f"100 % writes  fill_(1.0) float64           {best(lambda: f64.fill_(1.0), 4 * G):8.1f} GB/s")
print(f"100 % writes  fill_(1) bfloat16            {best(lambda: f64.view(torch.bfloat16).fill_(1.0), 4 * G):8.1f} GB/s")
print(f"100 % writes  zero_ (cudaMemsetAsync)      {best(lambda: f64.zero_(), 4 * G):8.1f} GB/s")
