"""Integer-tensor-core Gram (panel_tc.cu) vs the fp64 DMMA Gram on tall RGSQRF problems."""
import os, sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
torch.backends.cuda.matmul.allow_tf32 = False
def run(m, n, dist, i8):
    os.environ["LB_GRAM_I8"] = "1" if i8 else "0"
    ctx = qr.Context()
    g = torch.Generator(device="cuda").manual_seed(5)
    A0 = (torch.rand if dist == "uniform" else torch.randn)(m, n, device="cuda", generator=g)
    if dist == "graded":
        A0 = A0 * torch.logspace(0, -3, n, device="cuda")[None, :]
    A = qr.to_colmajor(A0); R = qr.colmajor_empty(n, n)
    ts = []
    for i in range(4):
        A.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); qr.later_rgsqrf(ctx, m, n, A, m, R, n); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    back = qr.backward_error(A0, A, R); orth = qr.orthogonality(A)
    out = (R.clone(), min(ts), back, orth, ctx.last_launch_count)
    ctx.close()
    return out
for (m, n) in ((65536, 256), (262144, 256), (131072, 1024), (1048576, 1024)):
    for dist in ("normal", "uniform", "graded"):
        R0, t0, b0, o0, l0 = run(m, n, dist, False)
        R1, t1, b1, o1, l1 = run(m, n, dist, True)
        d = ((R1 - R0).abs().max() / R0.abs().max()).item()
        print(f"{m}x{n} {dist:8s}: dmma {t0:8.3f} ms back {b0:.3e} orth/n {o0:.3e} | i8 {t1:8.3f} ms back {b1:.3e} "
              f"orth/n {o1:.3e} | max|dR|/max|R| {d:.2e} launches {l0}/{l1}", flush=True)
