import os, sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
os.environ["LB_APPLY_TC"] = "0"
ctx = qr.Context()
m = 16384
A = qr.colmajor_empty(m, 128); R = qr.colmajor_empty(128, 128)
for _ in range(2):
    A.normal_()
    qr.mgs_caqr_panel_256x128(ctx, m, 128, A, m, R, 128)
torch.cuda.synchronize()
