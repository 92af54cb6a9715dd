import os, sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
def run(m, n, minrows):
    os.environ["LB_GRAM_I8_MIN_ROWS"] = str(minrows)
    ctx = qr.Context()
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = torch.randn(m, n, device="cuda", generator=g)
    A = qr.to_colmajor(A0); R = qr.colmajor_empty(n, n)
    ts = []
    for i in range(5):
        A.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); qr.later_rgsqrf(ctx, m, n, A, m, R, n); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    b, o = qr.backward_error(A0, A, R), qr.orthogonality(A)
    ctx.close()
    return min(ts), b, o
for (m, n) in ((16384, 16384), (32768, 4096), (32768, 32768), (49152, 1024)):
    for mr in (65536, 16384):
        t, b, o = run(m, n, mr)
        print(f"{m}x{n} i8 from {mr:6d} rows: {t:8.3f} ms  back {b:.3e} orth/n {o:.3e}", flush=True)
