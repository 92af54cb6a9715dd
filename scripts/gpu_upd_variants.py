import os, sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
ctx = qr.Context()
m = 16384
for h in (1024, 2048, 4096, 8192):
    n = 2 * h
    Qh = qr.colmajor_empty(m, n, dtype=torch.float16); Qh.normal_()
    A = qr.colmajor_empty(m, n); A.normal_()
    Bh = qr.colmajor_empty(h, h, dtype=torch.float16); Bh.normal_()
    for var in (1, 2, 3):
        os.environ["LB_UPDATE_VARIANT"] = str(var)
        for _ in range(3): qr.gemm_update(ctx, Qh, 0, h, Bh, A[:, h:], Qh[:, h:], True)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); qr.gemm_update(ctx, Qh, 0, h, Bh, A[:, h:], Qh[:, h:], True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"h={h} variant={var}: {min(ts)*1e3:.1f} us  {2.0*h*h*m/min(ts)/1e9:.0f} TFLOPS", flush=True)
