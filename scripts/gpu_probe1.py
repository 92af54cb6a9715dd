"""First GPU bring-up: GEMM kernels vs torch, panel, full RGSQRF.  Run pieces via argv."""
import sys, time, json
import torch
sys.path.insert(0, '.')
from later_b200 import qr

torch.backends.cuda.matmul.allow_tf32 = False
dev = 'cuda'
ctx = qr.Context()

def cm(x):
    return qr.to_colmajor(x)

def t_gram(m, cols, colA, Mc, colB, Nc, splits):
    g = torch.Generator(device=dev).manual_seed(1)
    Q = cm(torch.randn(m, cols, device=dev, generator=g).half())
    Cout = qr.colmajor_empty(Mc, Nc)
    Ch = qr.colmajor_empty(Mc, Nc, dtype=torch.float16)
    Cout.fill_(-7); Ch.fill_(-7)
    qr.gemm_gram(ctx, Q, colA, Mc, colB, Nc, Cout, Ch, splits)
    torch.cuda.synchronize()
    ref = Q[:, colA:colA+Mc].float().t() @ Q[:, colB:colB+Nc].float()
    err = (Cout - ref).abs().max().item() / ref.abs().max().item()
    errh = (Ch.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"gram m={m} Mc={Mc} Nc={Nc} splits={splits}: rel err {err:.2e} fp16-copy {errh:.2e}", flush=True)
    return err

def t_update(m, K, Nc, sub=True):
    g = torch.Generator(device=dev).manual_seed(2)
    Q = cm(torch.randn(m, K + 64, device=dev, generator=g).half())
    B = cm((torch.randn(K, Nc, device=dev, generator=g) / 8).half())
    C0 = cm(torch.randn(m, Nc, device=dev, generator=g))
    Cc = cm(C0.clone())
    Ch = qr.colmajor_empty(m, Nc, dtype=torch.float16)
    qr.gemm_update(ctx, Q, 64, K, B, Cc, Ch, sub)
    torch.cuda.synchronize()
    prod = Q[:, 64:64+K].float() @ B.float()
    ref = C0 - prod if sub else prod
    err = (Cc - ref).abs().max().item() / ref.abs().max().item()
    errh = (Ch.float() - ref).abs().max().item() / ref.abs().max().item()
    print(f"update m={m} K={K} Nc={Nc} sub={sub}: rel err {err:.2e} fp16-copy {errh:.2e}", flush=True)
    return err

def t_panel(m):
    g = torch.Generator(device=dev).manual_seed(3)
    A0 = cm(torch.rand(m, 128, device=dev, generator=g))
    A = cm(A0.clone()); R = qr.colmajor_empty(128, 128); R.fill_(5)
    qr.mgs_caqr_panel_256x128(ctx, m, 128, A, m, R, 128)
    torch.cuda.synchronize()
    be = qr.backward_error(A0, A, R, torch.float64); orth = qr.orthogonality(A, torch.float64)
    low = torch.tril(R, -1).abs().max().item()
    print(f"panel m={m}: backward {be:.2e} orth {orth:.2e} lower-max {low:.1e} diagmin {R.diagonal().min().item():.3e}", flush=True)

def t_qr(m, n, dist='uniform', reps=3, graph=True):
    g = torch.Generator(device=dev).manual_seed(4)
    A0 = cm(torch.rand(m, n, device=dev, generator=g) if dist == 'uniform' else torch.randn(m, n, device=dev, generator=g))
    A = cm(A0.clone()); R = qr.colmajor_empty(n, n); R.fill_(3)
    c = qr.Context(use_graph=graph)
    times = []
    for r in range(reps):
        A.copy_(A0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        qr.later_rgsqrf(c, m, n, A, m, R, n)
        e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    be = qr.backward_error(A0, A, R); orth = qr.orthogonality(A)
    fl = 2.0*n*n*(m - n/3.0)
    print(f"rgsqrf {m}x{n} {dist} graph={graph}: times ms {['%.3f'%t for t in times]} -> {fl/min(times)/1e9:.1f} TFLOPS; backward {be:.3e} orth/n {orth:.3e} launches {c.last_launch_count}", flush=True)
    c.close()

what = sys.argv[1]
if what == 'gram':
    t_gram(1024, 512, 0, 128, 128, 128, 1)
    t_gram(1024, 512, 0, 256, 256, 256, 1)
    t_gram(4096, 512, 128, 128, 256, 128, 8)
    t_gram(16384, 2048, 0, 1024, 1024, 1024, 0)
    t_gram(8192, 8192, 0, 4096, 4096, 4096, 1)
elif what == 'update':
    t_update(1024, 128, 128)
    t_update(1024, 256, 256)
    t_update(4096, 512, 512)
    t_update(16384, 1024, 1024)
    t_update(2048, 256, 256, sub=False)
elif what == 'panel':
    for m in (256, 1024, 16384, 262144):
        t_panel(m)
elif what == 'qr':
    t_qr(1024, 1024); t_qr(1024, 1024, graph=False)
    t_qr(4096, 1024, 'normal')
    t_qr(16384, 1024, 'normal')
    t_qr(8192, 8192)
    t_qr(16384, 16384)
    t_qr(262144, 256, 'normal')
