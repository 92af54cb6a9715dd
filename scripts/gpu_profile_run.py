"""One un-graphed RGSQRF so that ncu sees every launch.  argv: m n [reps]"""
import sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
m, n = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
ctx = qr.Context(use_graph=False)
A = qr.colmajor_empty(m, n); R = qr.colmajor_empty(n, n)
for _ in range(reps):
    A.uniform_()
    qr.later_rgsqrf(ctx, m, n, A, m, R, n)
torch.cuda.synchronize()
print("done", ctx.last_launch_count)
