"""Sweep split-K factors for the small Gram products (run under ncu to get kernel durations)."""
import sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
ctx = qr.Context()
m = 16384
Qh = qr.colmajor_empty(m, 2048, dtype=torch.float16); Qh.normal_()
for h in (128, 256, 512, 1024):
    C = qr.colmajor_empty(h, h); Ch = qr.colmajor_empty(h, h, dtype=torch.float16)
    for s in (1, 2, 4, 8, 16, 32, 64, 128):
        for _ in range(2):
            qr.gemm_gram(ctx, Qh, 0, h, h, h, C, Ch, s)
        print("CFG", h, s, flush=True)
torch.cuda.synchronize()
