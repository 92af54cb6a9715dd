"""Two contexts on two streams of ONE device factor matrices whose fused node kernels (one CTA per 128-row tile,
grid-wide barriers) cannot both be resident at once: with cooperative launches the driver runs the two grids one
after the other; results must equal the serial ones.  Always run under `timeout`."""
import sys
import torch
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from later_b200 import qr

m, n, reps = 16384, 512, 6
g = torch.Generator(device="cuda").manual_seed(7)
A0 = [torch.randn(m, n, device="cuda", generator=g) for _ in range(2)]
torch.cuda.synchronize()
streams = [torch.cuda.Stream(), torch.cuda.Stream()]
ctxs = [qr.Context(stream=s) for s in streams]
ref = []
for i in range(2):                                   # serial reference
    with torch.cuda.stream(streams[i]):
        A = qr.to_colmajor(A0[i]); R = qr.colmajor_empty(n, n)
        qr.later_rgsqrf(ctxs[i], m, n, A, m, R, n)
    torch.cuda.synchronize()
    ref.append((A.clone(), R.clone()))
outs = [[], []]
for r in range(reps):                                # interleaved, nothing waits for anything
    for i in range(2):
        with torch.cuda.stream(streams[i]):
            A = qr.to_colmajor(A0[i]); R = qr.colmajor_empty(n, n)
            qr.later_rgsqrf(ctxs[i], m, n, A, m, R, n)
            outs[i].append((A, R))
torch.cuda.synchronize()
ok = all(torch.equal(A, ref[i][0]) and torch.equal(R, ref[i][1]) for i in range(2) for A, R in outs[i])
print("concurrent contexts:", "ok" if ok else "MISMATCH", flush=True)
sys.exit(0 if ok else 1)
