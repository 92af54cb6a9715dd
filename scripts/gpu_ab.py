"""A/B timing of whole factorisations under the current environment knobs (LB_*), one line per shape:
   LB_CHOL=1 python scripts/gpu_ab.py 16384x16384 1048576x1024"""
import sys
import torch
sys.path.insert(0, '.')
from later_b200 import qr

shapes = [tuple(map(int, a.split("x"))) for a in sys.argv[1:]] or [(16384, 16384), (262144, 256), (131072, 1024),
                                                                   (8192, 1024), (1048576, 1024)]
ctx = qr.Context()
g = torch.Generator(device="cuda").manual_seed(3000)
for m, n in shapes:
    A0 = torch.empty((n, m), device="cuda").normal_(generator=g).t()
    A = torch.empty((n, m), device="cuda").t()
    R = torch.zeros((n, n), device="cuda").t()
    ts = []
    for i in range(8):
        A.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); qr.later_rgsqrf(ctx, m, n, A, m, R, n); e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    back = qr.backward_error(A0, A, R) if m * n <= 2 ** 28 else float("nan")
    print(f"{m}x{n}: median {ts[len(ts) // 2]:.3f} ms  min {ts[0]:.3f}  launches {ctx.last_launch_count}  "
          f"backward {back:.2e}  info {ctx.last_info()}", flush=True)
    del A0, A, R
