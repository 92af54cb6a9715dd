"""Small shapes through every entry point, for compute-sanitizer (memcheck / racecheck / synccheck):
  compute-sanitizer --tool racecheck python scripts/gpu_sanitize_run.py
Shapes are small because the sanitizer slows kernels down by orders of magnitude; `tall` adds the
tensor-core panel kernels (integer Gram, tensor-core apply, cast-fused Gram) on 65536 rows."""
import sys
import torch
sys.path.insert(0, '.')
from later_b200 import qr

which = sys.argv[1:] or ["square", "panel", "panel32", "host", "ormqr"]
ctx = qr.Context(use_graph=False)
g = torch.Generator(device="cuda").manual_seed(1)


def check(name, A0, Q, R):
    back = qr.backward_error(A0, Q, R)
    orth = qr.orthogonality(Q)
    print(f"{name}: backward {back:.2e} orth {orth:.2e} info {ctx.last_info()}", flush=True)
    assert back < 1e-3 and orth < 1e-3


if "square" in which:
    A0 = torch.randn(1024, 512, device="cuda", generator=g)
    A = qr.to_colmajor(A0); R = qr.colmajor_empty(512, 512)
    qr.later_rgsqrf(ctx, 1024, 512, A, 1024, R, 512)
    check("rgsqrf 1024x512", A0, A, R)
if "panel" in which:
    A0 = torch.randn(2048, 128, device="cuda", generator=g)
    A = qr.to_colmajor(A0); R = qr.colmajor_empty(128, 128)
    qr.mgs_caqr_panel_256x128(ctx, 2048, 128, A, 2048, R, 128)
    check("panel 2048x128", A0, A, R)
if "panel32" in which:
    A0 = torch.randn(1000, 32, device="cuda", generator=g)
    A = qr.to_colmajor(A0); R = qr.colmajor_empty(32, 32)
    qr.mgs_caqr_panel_256x32(ctx, 1000, 32, A, 1000, R, 32)
    check("panel32 1000x32", A0, A, R)
if "host" in which:
    A0 = torch.randn(1024, 256, generator=torch.Generator().manual_seed(2))
    hA = torch.empty((256, 1024)).pin_memory().t(); hA.copy_(A0)
    hR = torch.zeros((256, 256)).pin_memory().t()
    qr.later_rgsqrf_host(ctx, 1024, 256, hA, 1024, hR, 256)
    check("rgsqrf_host 1024x256", A0.cuda(), hA.cuda(), torch.triu(hR.cuda()))
if "ormqr" in which:
    Y = torch.tril(torch.randn(512, 256, device="cuda", generator=g) * 0.1, -1); Y.diagonal().fill_(1.0)
    W = qr.to_colmajor(torch.randn(512, 256, device="cuda", generator=g) * 0.1)
    qr.later_ormqr(512, 256, W, 512, qr.to_colmajor(Y), 512, ctxt=ctx)
    torch.cuda.synchronize()
    print("ormqr 512x256 done", flush=True)
if "tall" in which:
    A0 = torch.randn(65536, 256, device="cuda", generator=g)
    A = qr.to_colmajor(A0); R = qr.colmajor_empty(256, 256)
    qr.later_rgsqrf(ctx, 65536, 256, A, 65536, R, 256)
    check("rgsqrf 65536x256 (tensor-core panel kernels)", A0, A, R)
if "qdwh" in which:
    H = torch.rand(256, 256, device="cuda", generator=g); H = 0.5 * (H + H.t())
    B = qr.colmajor_empty(512, 256)
    it = qr.later_qdwh_polar(ctx, 256, B, 512, None, 256, qr.to_colmajor(H))
    print("qdwh 256:", it, "iterations", flush=True)
torch.cuda.synchronize()
print("sanitize run done", flush=True)
