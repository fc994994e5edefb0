#!/usr/bin/env python
"""pack_book / unpack_book on a 2^18 x 253 float64 data set (the bench's shape), a few calls each: for an ncu DRAM-byte capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cantorrl_b200 import _lib, sim  # noqa: E402

dev = torch.device("cuda", 0)
n, T = 1 << 18, 252
book = sim.generate_paths_and_options(n, n_steps=T, model="gbm", device=dev)
L = _lib.lib()
sp = torch.cuda.current_stream(dev).cuda_stream
arrs = [torch.empty((n, T + 1), dtype=torch.float64, device=dev), torch.empty((n, T + 1), dtype=torch.float64, device=dev),
        torch.empty((n, T), dtype=torch.float64, device=dev), torch.empty((n, T), dtype=torch.float64, device=dev)]
ptrs = [a.data_ptr() for a in arrs]
for _ in range(3):
    _lib.check(L.cantor_unpack_book(book.tensor.data_ptr(), book.ld, n, T, _lib.F64, *ptrs, sp), "unpack")
    _lib.check(L.cantor_pack_book(*ptrs, _lib.F64, n, T, book.tensor.data_ptr(), book.ld, sp), "pack")
torch.cuda.synchronize()
print("algorithmic bytes per call: %.1f MB" % (48.0 * n * (T + 1) / 1e6))
