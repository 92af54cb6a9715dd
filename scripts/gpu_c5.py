"""Config 5 of BASELINE.json: RGSQRF 32768 x 32768 and explicit-Q formation (later_ormqr) on one B200,
this library and the reference build (oracle/_ref/libref_later.so) on the same inputs.
Writes gpurun_out/c5.json."""
import ctypes as C, json, sys, time
from pathlib import Path
import torch
sys.path.insert(0, '.')
from later_b200 import qr

ROOT = Path(__file__).resolve().parent.parent
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n = m = int(args[0]) if args else 32768
ORMQR_ONLY = "--ormqr-only" in sys.argv
torch.backends.cuda.matmul.allow_tf32 = False
out = {"shape": f"{m}x{n}", "note": "B200, warm, CUDA events, N(0,1) input, same device buffer for both"}

def ev_time(fn, restore, reps=3, warm=2):
    ts = []
    for i in range(warm + reps):
        restore()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]

def metrics(A0, Q, R):
    Rt = torch.triu(R)
    back = (torch.linalg.matrix_norm(Q @ Rt - A0) / torch.linalg.matrix_norm(A0)).item()
    G = Q.t() @ Q
    G.diagonal().sub_(1.0)
    return back, (torch.linalg.matrix_norm(G) / n).item()

ref = None
p = ROOT / "oracle" / "_ref" / "libref_later.so"
if p.exists():
    ref = C.CDLL(str(p))
    vp, ci = C.c_void_p, C.c_int
    ref.ref_later_rgsqrf.argtypes = [ci, ci, vp, ci, vp, ci, vp, ci, vp, ci]
    ref.ref_later_ormqr.argtypes = [ci, ci, vp, ci, vp, ci, vp]

gen = torch.Generator(device="cuda").manual_seed(3000)
if not ORMQR_ONLY:
    flops = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3
    ctx = qr.Context()
    A0 = torch.empty((n, m), device="cuda").normal_(generator=gen).t()
    A = torch.empty((n, m), device="cuda").t()
    R = torch.zeros((n, n), device="cuda").t()
    t = ev_time(lambda: qr.later_rgsqrf(ctx, m, n, A, m, R, n), lambda: A.copy_(A0))
    b, o = metrics(A0, A, R)
    out["rgsqrf_b200"] = {"ms": t, "tflops": flops / t / 1e9, "backward": b, "orth_over_n": o,
                          "launches": ctx.last_launch_count}
    print(out["rgsqrf_b200"], flush=True)
    if ref:
        work = torch.zeros(m // 256 * 32 * n + (1 << 20), device="cuda")
        hwork = torch.zeros(m * n, device="cuda", dtype=torch.float16)
        def run_ref():
            rc = ref.ref_later_rgsqrf(m, n, A.data_ptr(), m, R.data_ptr(), n, work.data_ptr(), work.numel(),
                                      hwork.data_ptr(), hwork.numel())
            assert rc == 0
        R.zero_()
        t = ev_time(run_ref, lambda: A.copy_(A0), reps=2, warm=1)
        b, o = metrics(A0, A, R)
        out["rgsqrf_reference"] = {"ms": t, "tflops": flops / t / 1e9, "backward": b, "orth_over_n": o}
        print(out["rgsqrf_reference"], flush=True)
        del work, hwork
    del A0, A, R
    ctx.close()
    torch.cuda.empty_cache()


# ---- explicit Q from a WY pair (later_ormqr): Y unit lower trapezoidal, W of the same scale
ctx = qr.Context()
Y0 = torch.empty((n, m), device="cuda").normal_(generator=gen).mul_(0.01).t()     # column-major m x n
Y0 = torch.tril(Y0, -1); Y0.diagonal().fill_(1.0)
Y0 = Y0.t().contiguous().t()
W0 = torch.empty((n, m), device="cuda").normal_(generator=gen).mul_(0.01).t()
W = torch.empty((n, m), device="cuda").t()
h = n // 2
# fp64 check on every 64th row (rows are independent once T = Y1^T W2 is known)
T = (Y0[:, :h].double().t() @ W0[:, h:].double())
rows = torch.arange(0, m, 64, device="cuda")
Ws = W0[rows].double()
Ws[:, h:] -= Ws[:, :h] @ T
ref64 = -(Ws @ Y0[:n, :n].double().t())
ref64[torch.arange(rows.numel(), device="cuda"), rows] += 1.0
del T, Ws
scale = ref64.abs().max().item()
oflops = 2.0 * h * h * m * 2 + 2.0 * m * n * n
t = ev_time(lambda: qr.later_ormqr(m, n, W, m, Y0, m, ctxt=ctx), lambda: W.copy_(W0))
err = ((W[rows].double() - ref64).abs().max().item()) / scale
out["ormqr_b200"] = {"ms": t, "tflops_executed": oflops / t / 1e9, "max_err_vs_fp64": err,
                     "launches": ctx.last_launch_count}
print(out["ormqr_b200"], flush=True)
if ref:
    work = torch.zeros(m * n, device="cuda")
    def run_ref_o():
        rc = ref.ref_later_ormqr(m, n, W.data_ptr(), m, Y0.data_ptr(), m, work.data_ptr())
        assert rc == 0
    t = ev_time(run_ref_o, lambda: W.copy_(W0), reps=2, warm=1)
    err = ((W[rows].double() - ref64).abs().max().item()) / scale
    out["ormqr_reference"] = {"ms": t, "tflops_executed": oflops / t / 1e9, "max_err_vs_fp64": err}
    print(out["ormqr_reference"], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/c5.json").write_text(json.dumps(out, indent=1))
