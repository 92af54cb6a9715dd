"""Run one building block a few times (for ncu --set full captures).  argv: what m h"""
import sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
what, m, h = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
ctx = qr.Context()
if what in ('update', 'gram'):
    n = 2 * h
    Qh = qr.colmajor_empty(m, n, dtype=torch.float16); Qh.normal_()
    A = qr.colmajor_empty(m, n); A.normal_()
    if what == 'gram':
        C = qr.colmajor_empty(h, h); Ch = qr.colmajor_empty(h, h, dtype=torch.float16)
        for _ in range(3): qr.gemm_gram(ctx, Qh, 0, h, h, h, C, Ch, 0)
    else:
        Bh = qr.colmajor_empty(h, h, dtype=torch.float16); Bh.normal_()
        for _ in range(3): qr.gemm_update(ctx, Qh, 0, h, Bh, A[:, h:], Qh[:, h:], True)
elif what == 'panel':
    P = qr.colmajor_empty(m, 128); R = qr.colmajor_empty(128, 128)
    for _ in range(3):
        P.uniform_()
        qr.mgs_caqr_panel_256x128(ctx, m, 128, P, m, R, 128)
torch.cuda.synchronize()
print('ok')
