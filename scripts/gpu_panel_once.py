import os, sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
m = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
os.environ["LB_APPLY_TC"] = "1"
ctx = qr.Context()
A = qr.colmajor_empty(m, 128); R = qr.colmajor_empty(128, 128)
for _ in range(reps):
    A.normal_()
    qr.mgs_caqr_panel_256x128(ctx, m, 128, A, m, R, 128)
torch.cuda.synchronize()
