"""numpy model of the ROW-SHARDED recursion (later_b200_rgsqrf_dist, later_b200/csrc/rgsqrf.cu): every rank
runs the recursion on its own rows; the only exchanges are the sum over the ranks of each panel's Gram matrix
(fp64) and of each R12 block (fp32).  Test infrastructure: it restates this repository's multi-GPU algorithm,
rounding by rounding, so that the gloo test can check on CPU what the algorithm promises - the same R on every
rank, the accuracy of the single-GPU factorisation - without a GPU.  `allreduce(array) -> array` is injected
(identity for one rank)."""
import numpy as np
import scipy.linalg

PANEL = 128


def _h(a):
    return a.astype(np.float16).astype(np.float32)


def _panel(Ap, allreduce):
    """Gram-matrix panel (panel.cu): G = sum over ranks of A_p^T A_p with exact products and fp64 sums,
    R = chol(G) in fp64, Q_p = A_p R^-1 by substitution in fp32."""
    G = allreduce(Ap.astype(np.float64).T @ Ap.astype(np.float64))
    R = np.linalg.cholesky(G).T
    R32 = R.astype(np.float32)
    Q = scipy.linalg.solve_triangular(R32.T, Ap.T.astype(np.float32), lower=True).T.astype(np.float32)
    return Q, R32


def _qr(Ap, R, c0, w, allreduce):
    if w <= PANEL:
        Q, r = _panel(Ap[:, c0:c0 + w], allreduce)
        Ap[:, c0:c0 + w] = Q
        R[c0:c0 + w, c0:c0 + w] = r
        return
    h = w // 2
    _qr(Ap, R, c0, h, allreduce)
    Q1h = _h(Ap[:, c0:c0 + h])
    # this rank's share of R12 = Q1^T A2 (fp16 operands, fp32 accumulation), summed over the ranks in fp32
    R12 = allreduce((Q1h.T @ _h(Ap[:, c0 + h:c0 + w])).astype(np.float32))
    R[c0:c0 + h, c0 + h:c0 + w] = R12
    Ap[:, c0 + h:c0 + w] -= Q1h @ _h(R12)
    _qr(Ap, R, c0 + h, h, allreduce)


def rgsqrf_sharded(Ap: np.ndarray, allreduce=lambda x: x):
    """Q_p (this rank's rows of Q) and R (the same on every rank) of the row-sharded recursion."""
    Ap = np.array(Ap, dtype=np.float32, order="F")
    n = Ap.shape[1]
    R = np.zeros((n, n), dtype=np.float32)
    _qr(Ap, R, 0, n, allreduce)
    return Ap, R
