"""TEST INFRASTRUCTURE - CPU restatement (numpy) of the reference's RGSQRF path.

This is the checker, not the product: only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline leg may import it.  It restates, step for step, what the reference computes, with
the same precision at every step (fp32 everywhere, operands of the two trailing-update products
rounded to fp16 with round-to-nearest, fp32 accumulation):

    qr()                      QR/later_rgsqrf.cu:25-60     recursion, NMIN = 128
    s2h()                     util/util.cu:24-32           fp32 -> fp16 RN cast
    mgs_caqr_panel_256x128()  QR/panel.cu:10-63            128-col panel = 4 x 32-col CAQR + block GS
    mgs_caqr_panel_256x32()   QR/panel.cu:65-134           TSQR tree over 256-row blocks
    mgs_kernel2()             QR/panel.cu:246-325          256 x 32 modified Gram-Schmidt
    later_ormqr/_ormqr2()     QR/later_ormqr.cu:18-85      explicit Q from WY
    later_qdwh_polar()        EVD/later_qdwh_polar.cu:24-110  QDWH polar iteration, the caller of later_rgsqrf
    later_rhouqr()            QR/later_rhouqr.cu:21-277    recursive Householder QR in WY form (+ panel.cu:341-378)
    check_result/check_otho   test/test_qr.cu:216-268      the driver's self-consistency metrics

Third-party arithmetic: every GEMM on the path is cuBLAS (closed source; 12.9.1.4 in this image;
call sites QR/later_rgsqrf.cu:45-56, QR/panel.cu:21-61,97-130, QR/later_ormqr.cu:27-58).  Its
published semantics are "C = alpha op(A) op(B) + beta C with the stated input/compute types,
summation order unspecified"; the restatement uses numpy's sgemm with the same types.  Results
therefore agree with the reference up to fp32 summation order, not bit for bit.

Pinning: the reference has no golden vectors.  The oracle is pinned against
outputs of the reference itself, generated on a B200 from oracle/_ref/libref_later.so by
tests/golden/make_golden.py and committed under tests/golden/ (see tests/test_oracle.py).
"""
from __future__ import annotations

import numpy as np

NMIN = 128  # QR/later_rgsqrf.cu:23
F32 = np.float32


def s2h(a: np.ndarray) -> np.ndarray:
    """fp32 -> fp16 round-to-nearest-even (util/util.cu:24-32, __float2half)."""
    return np.asarray(a, dtype=F32).astype(np.float16)


def _gemm_f32(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """fp32 x fp32 -> fp32 product (cublasSgemm / GemmEx CUDA_R_32F)."""
    return np.matmul(np.asarray(a, dtype=F32), np.asarray(b, dtype=F32)).astype(F32)


def _gemm_tc(ah: np.ndarray, bh: np.ndarray) -> np.ndarray:
    """fp16 inputs, fp32 accumulate and output (cublasGemmEx CUDA_R_16F in, CUDA_R_32F compute)."""
    return np.matmul(ah.astype(F32), bh.astype(F32)).astype(F32)


def mgs_kernel2(A: np.ndarray) -> np.ndarray:
    """Right-looking MGS of one <=256 x 32 block, in place; returns R (32 x 32, zeros below the
    diagonal).  Order of operations as QR/panel.cu:273-313: normalise column k, then project it
    out of every later column."""
    mm, n = A.shape
    mnmin = min(mm, n)
    R = np.zeros((n, n), dtype=F32)
    for k in range(mnmin):
        nu = F32(np.dot(A[:, k], A[:, k]))
        normx = F32(np.sqrt(nu))
        R[k, k] = normx
        A[:, k] *= F32(1.0) / normx
        if k + 1 < n:
            r = (A[:, k][None, :] @ A[:, k + 1:]).astype(F32).ravel()   # q_k^T a_j
            A[:, k + 1:] -= np.outer(A[:, k], r).astype(F32)
            R[k, k + 1:] = r
    return R


def mgs_caqr_panel_256x32(A: np.ndarray) -> np.ndarray:
    """CAQR of an m x 32 panel, in place; returns R (32 x 32).  QR/panel.cu:65-134: MGS on each
    256-row block, R factors stacked (block b -> rows 32b..32b+31), recursion on the stack, then
    A_blk <- Q_blk * W_blk.  The m % 256 != 0 remainder block follows :105-132."""
    m, n = A.shape
    assert n == 32
    if m <= 256:
        return mgs_kernel2(A)
    nb = m // 256
    r = m % 256
    nblk = nb + (1 if r else 0)
    stack = np.zeros((nblk * 32, 32), dtype=F32)
    for b in range(nblk):
        blk = A[b * 256:min(m, (b + 1) * 256), :]
        stack[b * 32:(b + 1) * 32, :] = mgs_kernel2(blk)
    R = mgs_caqr_panel_256x32(stack)          # stack <- its own Q
    for b in range(nblk):
        rows = slice(b * 256, min(m, (b + 1) * 256))
        A[rows, :] = _gemm_f32(A[rows, :], stack[b * 32:(b + 1) * 32, :])
    return R


def mgs_caqr_panel_256x128(A: np.ndarray, R: np.ndarray) -> None:
    """128-column panel, in place; R is the 128 x 128 view to fill (QR/panel.cu:10-63)."""
    m, n = A.shape
    assert n == 128
    def gs(q, a, r_out):
        r = _gemm_f32(q.T, a)            # R12 = Q^T A     (cublasSgemm T,N)
        a -= _gemm_f32(q, r)             # A  -= Q R12     (cublasSgemm N,N)
        r_out[...] = r
    R[0:32, 0:32] = mgs_caqr_panel_256x32(A[:, 0:32])
    gs(A[:, 0:32], A[:, 32:64], R[0:32, 32:64])
    R[32:64, 32:64] = mgs_caqr_panel_256x32(A[:, 32:64])
    gs(A[:, 0:64], A[:, 64:128], R[0:64, 64:128])
    R[64:96, 64:96] = mgs_caqr_panel_256x32(A[:, 64:96])
    gs(A[:, 64:96], A[:, 96:128], R[64:96, 96:128])
    R[96:128, 96:128] = mgs_caqr_panel_256x32(A[:, 96:128])


def _qr(A: np.ndarray, R: np.ndarray) -> None:
    """QR/later_rgsqrf.cu:25-60 on column views A (m x w) and R (w x w)."""
    m, w = A.shape
    if w <= NMIN:
        mgs_caqr_panel_256x128(A, R)
        return
    h = w // 2
    _qr(A[:, :h], R[:h, :h])
    ah = s2h(A[:, :h])                       # Q1 -> fp16           (:43)
    bh = s2h(A[:, h:])                       # A2 -> fp16           (:44)
    r12 = _gemm_tc(ah.T, bh)                 # R12 = Q1^T A2        (:45-49)
    R[:h, h:] = r12
    r12h = s2h(r12)                          # R12 -> fp16          (:50-51)
    A[:, h:] -= _gemm_tc(ah, r12h)           # A2 -= Q1 R12         (:52-56)
    _qr(A[:, h:], R[h:, h:])


def later_rgsqrf(A: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Returns (Q, R) for an m x n fp32 matrix, n = 128 * 2^k, m >= n (QR/later_rgsqrf.cu:62-79).
    Unlike the reference, R's never-written lower blocks are returned as zeros."""
    A = np.array(A, dtype=F32, order="F", copy=True)
    m, n = A.shape
    if n % NMIN or (n // NMIN) & (n // NMIN - 1):
        raise ValueError("n must be 128 * 2^k")
    if m < n:
        raise ValueError("m must be >= n")
    R = np.zeros((n, n), dtype=F32, order="F")
    _qr(A, R)
    return A, R


def later_ormqr2(W: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """W <- I(m x n) - W Y[0:n, 0:n]^T, all fp32 (QR/later_ormqr.cu:66-85)."""
    W = np.asarray(W, dtype=F32)
    m, n = W.shape
    out = np.eye(m, n, dtype=F32)
    out -= _gemm_f32(W, np.asarray(Y, dtype=F32)[:n, :n].T)
    return out


def later_ormqr(W: np.ndarray, Y: np.ndarray) -> np.ndarray:
    """work = Y1^T W2; W2 -= W1 work; then later_ormqr2 (QR/later_ormqr.cu:18-64)."""
    W = np.array(W, dtype=F32, copy=True)
    Y = np.asarray(Y, dtype=F32)
    m, n = W.shape
    h = n // 2
    work = _gemm_f32(Y[:, :h].T, W[:, h:])
    W[:, h:] -= _gemm_f32(W[:, :h], work)
    return later_ormqr2(W, Y)


def later_qdwh_polar(X0: np.ndarray, smin_est: float = 0.0002070391384, max_iter: int = 10):
    """QDWH polar iteration as EVD/later_qdwh_polar.cu:24-110 runs it: normalise by the Frobenius norm
    (:26-31), L = smin_est / sqrt(n) with the reference's hard-coded smin_est (:37-38), then per
    iteration the dynamically weighted Halley coefficients (:60-66), RGSQRF of the stacked
    [sqrt(c) X; I] (:71-79), X <- (a - b/c)/sqrt(c) * fp16(Q1) fp16(Q2)^T + (b/c) X (:86-98) and
    the symmetrisation (X + X^T)/2 (:100, generateNewU :10-21 - restated as the averaging it is
    meant to be: the kernel itself races between thread blocks).  Stops as :53-57 does.
    Returns (U, iterations)."""
    X = np.array(X0, dtype=F32, order="F", copy=True)
    n = X.shape[0]
    X *= F32(1.0) / F32(np.sqrt(np.sum(X.astype(np.float64) ** 2)))
    L = F32(smin_est) / F32(np.sqrt(F32(n)))
    eps = F32(2e-4)
    tol1 = F32(10.0) * eps / F32(2.0)
    tol3 = F32(tol1 ** (1.0 / 3.0))
    prev = None
    it = 0
    for it in range(max_iter):
        if it > 0:
            diff = np.sqrt(np.sum((prev.astype(np.float64) - X.astype(np.float64)) ** 2))
            if diff < tol3 and F32(1.0) - L < tol1:
                break
        L2 = F32(L * L)
        dd = F32((F32(4.0) * (1 - L2) / (L2 * L2)) ** (1.0 / 3.0))
        sqd = F32(np.sqrt(1 + dd))
        a = F32(sqd + np.sqrt(8 - 4 * dd + 8 * (2 - L2) / (L2 * sqd)) / 2)
        b = F32((a - 1) * (a - 1) / 4)
        c = F32(a + b - 1)
        # clamped at 1 (mathematically L <= 1; the reference's fp32 update can overshoot by an ulp, after
        # which its coefficients are NaN - the clamp is the one deliberate deviation of this restatement)
        L = F32(min(F32(L * (a + b * L2) / (1 + c * L2)), F32(1.0)))
        B = np.empty((2 * n, n), dtype=F32, order="F")
        B[:n] = X * F32(np.sqrt(c))
        B[n:] = np.eye(n, dtype=F32)
        Q, _ = later_rgsqrf(B)
        W = _gemm_tc(s2h(Q[:n]), s2h(Q[n:]).T) * F32((a - b / c) / np.sqrt(c)) + F32(b / c) * X
        prev = X
        X = (F32(0.5) * (W + W.T)).astype(F32)
    else:
        it = max_iter
    return X, it


def _hou_caqr_panel(A: np.ndarray) -> np.ndarray:
    """hou_caqr_panel<256,32> (QR/panel.cu:341-378): Householder QR of every 256-row block, recursion on the
    stacked R factors, Q_blk <- Q_blk W_blk.  In place; returns R.  The block kernel (hou_kernel3,
    QR/panel.cu:386-558) is a textbook Householder QR that leaves the explicit Q; LAPACK's stands in for it."""
    m, n = A.shape
    if m <= 256:
        q, r = np.linalg.qr(A.astype(F32))
        A[...] = q.astype(F32)
        return r.astype(F32)
    nblk = (m + 255) // 256
    stack = np.zeros((nblk * n, n), dtype=F32)
    for b in range(nblk):
        blk = A[b * 256:min(m, (b + 1) * 256), :]
        q, r = np.linalg.qr(blk.astype(F32))
        blk[...] = q.astype(F32)
        stack[b * n:b * n + r.shape[0], :] = r
    R = _hou_caqr_panel(stack)
    for b in range(nblk):
        rows = slice(b * 256, min(m, (b + 1) * 256))
        A[rows, :] = _gemm_f32(A[rows, :], stack[b * n:(b + 1) * n, :])
    return R


def _rhouqr(A: np.ndarray, W: np.ndarray, R: np.ndarray, top: bool, merge_top: bool) -> None:
    """qr() of QR/later_rhouqr.cu:54-233 on views A, W (m x n) and R (n x n)."""
    m, n = A.shape
    if n <= 32:
        R[...] = np.triu(_hou_caqr_panel(A))                       # A <- Q               (:58)
        V = np.eye(m, n, dtype=F32) - A                            # I - Q                (:61-63)
        # reconstructY (:239-277): LU without pivoting of the top block, Y = [L; (I-Q)_2 U^-1]
        M = V[:n].astype(F32).copy()
        L = np.eye(n, dtype=F32)
        for j in range(n):
            L[j + 1:, j] = M[j + 1:, j] / M[j, j]
            M[j + 1:, j:] -= np.outer(L[j + 1:, j], M[j, j:]).astype(F32)
        U = np.triu(M)
        A[:n] = L
        A[n:] = np.linalg.solve(U.T.astype(np.float64), V[n:].T.astype(np.float64)).T.astype(F32)
        W[...] = np.linalg.solve(L.astype(np.float64), V.T.astype(np.float64)).T.astype(F32)   # (I-Q) L^-T (:67-74)
        return
    h = n // 2
    _rhouqr(A[:, :h], W[:, :h], R[:h, :h], False, merge_top)
    mul = _gemm_f32 if (h <= 128 or m <= 128) else (lambda a, b: _gemm_tc(s2h(a), s2h(b)))   # (:83, :104-137)
    A[:, h:] -= mul(A[:, :h], mul(W[:, :h].T, A[:, h:]))           # A2 <- Q1^T A2
    _rhouqr(A[h:, h:], W[h:, h:], R[h:, h:], False, merge_top)
    R[:h, h:] = A[:h, h:]                                          # (:152-154)
    A[:h, h:] = 0
    if not top or merge_top:       # (the guard the reference has commented out at :165; its driver's
        W[:, h:] -= mul(W[:, :h], mul(A[:, :h].T, W[:, h:]))       # later_ormqr performs the top merge)


def later_rhouqr(A: np.ndarray, merge_top: bool = False):
    """Returns (Y, W, R): A = (I - W Y^T) R, Y unit lower trapezoidal (QR/later_rhouqr.cu:26-41)."""
    Y = np.array(A, dtype=F32, order="F", copy=True)
    m, n = Y.shape
    W = np.zeros((m, n), dtype=F32, order="F")
    R = np.zeros((n, n), dtype=F32, order="F")
    _rhouqr(Y, W, R, True, merge_top)
    return Y, W, R


def check_result(A: np.ndarray, Q: np.ndarray, R: np.ndarray) -> float:
    """||A - Q R||_F / ||A||_F (test/test_qr.cu:216-228), evaluated in fp64."""
    A = np.asarray(A, dtype=np.float64)
    return float(np.linalg.norm(A - np.asarray(Q, np.float64) @ np.asarray(R, np.float64)) /
                 np.linalg.norm(A))


def check_otho(Q: np.ndarray) -> float:
    """||I - Q^T Q||_F / n - note the division by n (test/test_qr.cu:245-268)."""
    Q = np.asarray(Q, dtype=np.float64)
    n = Q.shape[1]
    return float(np.linalg.norm(np.eye(n) - Q.T @ Q) / n)


def lapack_qr(A: np.ndarray) -> tuple[np.ndarray, np.ndarray]:
    """Host LAPACK sgeqrf + sorgqr with explicit Q and the sign convention r_ii > 0 (the CPU
    stand-in BASELINE.json names: the reference has no CPU path)."""
    from scipy.linalg import lapack
    a = np.array(A, dtype=F32, order="F", copy=True)
    m, n = a.shape
    qr_, tau, _, info = lapack.sgeqrf(a, overwrite_a=1)
    if info:
        raise RuntimeError(f"sgeqrf info={info}")
    R = np.triu(qr_[:n, :]).astype(F32)
    Q, _, info = lapack.sorgqr(qr_[:, :n], tau, overwrite_a=1)
    if info:
        raise RuntimeError(f"sorgqr info={info}")
    s = np.sign(np.diag(R)).astype(F32)
    s[s == 0] = 1
    return (Q * s[None, :]).astype(F32), (R * s[:, None]).astype(F32)
