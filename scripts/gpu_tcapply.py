"""Tensor-core apply vs forward-substitution apply on the same panels: agreement, orthogonality, time."""
import os, sys, torch
sys.path.insert(0, '.')
from later_b200 import qr
ctx = qr.Context()
qr.lib.later_b200_set_graph(ctx._h, 0) if hasattr(qr, "lib") else None
def run(m, dist, tc):
    os.environ["LB_APPLY_TC"] = "1" if tc else "0"
    g = torch.Generator(device="cuda").manual_seed(7)
    A0 = (torch.rand if dist == "uniform" else torch.randn)(m, 128, device="cuda", generator=g)
    if dist == "graded":
        A0 = A0 * torch.logspace(0, -4, 128, device="cuda")[None, :]
    A = qr.to_colmajor(A0); R = qr.colmajor_empty(128, 128)
    ts = []
    for i in range(4):
        A.copy_(A0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); qr.mgs_caqr_panel_256x128(ctx, m, 128, A, m, R, 128); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    Q = A.double(); Rd = torch.triu(R).double()
    orth = torch.linalg.matrix_norm(Q.t() @ Q - torch.eye(128, device="cuda", dtype=torch.float64)).item()
    back = (torch.linalg.matrix_norm(Q @ Rd - A0.double()) / torch.linalg.matrix_norm(A0.double())).item()
    return A.clone(), min(ts), orth, back
for m in (1000, 16384, 131072, 262144, 1048576):
    for dist in ("normal", "uniform", "graded"):
        Q0, t0, o0, b0 = run(m, dist, False)
        Q1, t1, o1, b1 = run(m, dist, True)
        d = ((Q1 - Q0).abs().max() / Q0.abs().max()).item()
        print(f"m={m:8d} {dist:8s}: fwd-subst {t0*1e3:8.1f} us orth {o0:.2e} back {b0:.2e} | tensor-core {t1*1e3:8.1f} us "
              f"orth {o1:.2e} back {b1:.2e} | max|dQ|/max|Q| {d:.2e}", flush=True)
