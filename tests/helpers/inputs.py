"""Seeded test inputs and the reference driver's two metrics on the GPU (torch supplies memory and
fp64 matmuls for the CHECK only)."""
from __future__ import annotations

import math

import torch


def cond_matrix(m: int, n: int, kappa: float, seed: int, device="cuda") -> torch.Tensor:
    """m x n fp32 with singular values log-spaced between 1 and 1/kappa: A = G diag(sigma) V^T,
    G Gaussian / sqrt(m) (nearly orthonormal columns for m >> n), V orthogonal."""
    g = torch.Generator(device=device).manual_seed(seed)
    G = torch.randn(m, n, device=device, generator=g, dtype=torch.float32) / math.sqrt(m)
    V = torch.linalg.qr(torch.randn(n, n, device=device, generator=g, dtype=torch.float64))[0]
    sig = torch.logspace(0.0, -math.log10(kappa), n, device=device, dtype=torch.float64)
    out = torch.empty(m, n, device=device, dtype=torch.float32)
    step = max(1, (1 << 27) // n)               # bounded fp64 temporaries
    for r0 in range(0, m, step):
        out[r0:r0 + step] = ((G[r0:r0 + step].double() * sig) @ V.t()).float()
    return out


def metrics(A0: torch.Tensor, Q: torch.Tensor, R: torch.Tensor) -> tuple[float, float]:
    """(||A - Q R||_F / ||A||_F, ||I - Q^T Q||_F / n) evaluated in fp64
    (checkResult / checkOtho, reference test/test_qr.cu:216-268)."""
    m, n = A0.shape
    Rd = torch.triu(R.double())
    res2 = 0.0
    G = torch.zeros(n, n, device=A0.device, dtype=torch.float64)
    step = max(1, (1 << 26) // n)
    for r0 in range(0, m, step):
        q = Q[r0:r0 + step].double()
        res2 += float(torch.linalg.norm(A0[r0:r0 + step].double() - q @ Rd) ** 2)
        G += q.t() @ q
    G.diagonal().sub_(1.0)
    back = math.sqrt(res2) / float(torch.linalg.norm(A0.double()))
    return back, float(torch.linalg.norm(G)) / n
