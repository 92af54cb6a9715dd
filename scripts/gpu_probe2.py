"""Per-kernel timing of the RGSQRF building blocks at the 16384^2 shapes (CUDA events, warm)."""
import sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
dev = 'cuda'
ctx = qr.Context()
m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)

Qh = qr.colmajor_empty(m, n, dtype=torch.float16); Qh.normal_()
A = qr.colmajor_empty(m, n); A.normal_()
h = n // 2
while h >= 128:
    C = qr.colmajor_empty(h, h); Ch = qr.colmajor_empty(h, h, dtype=torch.float16)
    t = timeit(lambda: qr.gemm_gram(ctx, Qh, 0, h, h, h, C, Ch, 0))
    fl = 2.0 * h * h * m
    print(f"gram   h={h:5d}: {t*1e3:9.1f} us  {fl/t/1e9:8.1f} TFLOPS  launches {ctx.last_launch_count}", flush=True)
    Bh = qr.colmajor_empty(h, h, dtype=torch.float16); Bh.normal_()
    Cv = A[:, h:2*h]
    Chv = Qh[:, h:2*h]
    t = timeit(lambda: qr.gemm_update(ctx, Qh, 0, h, Bh, Cv, Chv, True))
    by = m * h * (2 + 4 + 4 + 2) + h * h * 2
    print(f"update h={h:5d}: {t*1e3:9.1f} us  {fl/t/1e9:8.1f} TFLOPS  {by/t/1e6:8.1f} GB/s", flush=True)
    h //= 2
P = qr.colmajor_empty(m, 128); R = qr.colmajor_empty(128, 128)
def panel():
    P.uniform_()
    qr.mgs_caqr_panel_256x128(ctx, m, 128, P, m, R, 128)
def fill():
    P.uniform_()
tp = timeit(panel); tf = timeit(fill)
print(f"panel m={m}: {(tp-tf)*1e3:.1f} us (4 kernels)  -> {8.0*m*128/(tp-tf)/1e6:.1f} GB/s algorithmic", flush=True)
