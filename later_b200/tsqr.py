"""Row-sharded tall-skinny QR over several GPUs (one process per GPU, torch.distributed).

The reference has no multi-GPU code; the structure is the one its CAQR panel already uses inside
one GPU across 256-row blocks (reference QR/panel.cu:87-104), lifted to ranks:

  1. every rank factors its own row block   A_p = Q_p R_p          (later_rgsqrf, local)
  2. the P small R_p (n x n) are exchanged                          (one all-gather: P * 4n^2 bytes)
  3. every rank factors the same stack [R_0; ...; R_{P-1}] = W R    (redundantly; identical bits
     on every rank because inputs and code are identical, so no broadcast of R or W is needed)
  4. Q_p <- Q_p W_p                                                 (tensor-core GEMM, local)

Only step 2 communicates.  `local_qr`, `stack_qr` and `apply_w` are injection points so that the
host-side logic can be exercised on CPU (gloo) with the numpy oracle standing in for the kernels.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist


_default_ctxs: dict = {}


def stack_from_gathered(gathered_t: torch.Tensor) -> torch.Tensor:
    """gathered_t[p] holds R_p^T row-major (= R_p column-major), shape (P, n, n).  Returns the
    column-major (P*n) x n stack [R_0; R_1; ...] as a tensor of shape (P*n, n), strides (1, P*n)."""
    P, n, _ = gathered_t.shape
    st = gathered_t.permute(1, 0, 2).reshape(n, P * n).contiguous()   # st[j, p*n + i] = R_p[i, j]
    return st.t()


def tsqr_rgsqrf(m_local: int, n: int, A: torch.Tensor, lda: int, R: torch.Tensor, ldr: int,
                group=None,
                local_qr: Callable | None = None,
                stack_qr: Callable | None = None,
                apply_w: Callable | None = None,
                ctxs=None, host_A: torch.Tensor | None = None) -> None:
    """In place: A (this rank's m_local x n row block, column-major) <- its block of the global Q;
    R (n x n) <- the global R factor (identical on every rank).  With `host_A` (host tensor of the
    same shape, ideally pinned) the row block is taken from there instead: it crosses PCIe while the
    local factorisation is already working on the columns that have arrived, and A is output only."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0

    if local_qr is not None and host_A is not None:
        raise ValueError("host_A needs the built-in local_qr")
    if local_qr is None:
        from . import qr as _qr
        if ctxs is None:                      # one pair per device, kept: graphs and workspaces are reused
            dev = torch.cuda.current_device()
            if dev not in _default_ctxs:
                _default_ctxs[dev] = (_qr.Context(dev), _qr.Context(dev))
            ctxs = _default_ctxs[dev]
        main_ctx, stack_ctx = ctxs

        def local_qr(m, n_, a, lda_, r, ldr_):
            if host_A is not None:
                _qr.later_rgsqrf_stream_in(main_ctx, m, n_, host_A, host_A.stride(1), a, lda_, r, ldr_)
            else:
                _qr.later_rgsqrf(main_ctx, m, n_, a, lda_, r, ldr_)

        def stack_qr(m, n_, s, lds, r, ldr_):
            _qr.later_rgsqrf(stack_ctx, m, n_, s, lds, r, ldr_)

        def apply_w(m, n_, q, ldq, w, ldw):
            _qr.tsqr_apply(main_ctx, m, n_, q, ldq, w, ldw)

    if world == 1:
        local_qr(m_local, n, A, lda, R, ldr)
        return

    # 1. local factorisation into a contiguous n x n R_p
    Rp = torch.empty((n, n), device=A.device, dtype=A.dtype).t()        # column-major, ld = n
    local_qr(m_local, n, A, lda, Rp, n)

    # 2. exchange: all-gather the column-major storage of every R_p
    gathered = torch.empty((world * n, n), device=A.device, dtype=A.dtype)
    dist.all_gather_into_tensor(gathered, Rp.t().contiguous(), group=group)
    gathered = gathered.view(world, n, n)

    # 3. redundant QR of the stack (canonical order: rank 0 on top)
    S = stack_from_gathered(gathered)                                    # (P*n) x n, ld = P*n
    Rs = torch.empty((n, n), device=A.device, dtype=A.dtype).t()
    stack_qr(world * n, n, S, world * n, Rs, n)

    # 4. back-multiplication with this rank's n x n block of the stack's Q
    W = S[rank * n:(rank + 1) * n, :]                                    # column-major view, ld = P*n
    apply_w(m_local, n, A, lda, W, world * n)
    R[:n, :n].copy_(Rs)


class MultiGpu:
    """Single-process driver of the same algorithm through the C ABI (later_b200_tsqr_mgpu): one host
    thread, P devices, one NCCL all-gather of the local R factors.  A[p], R[p] live on device p."""

    def __init__(self, devices):
        import ctypes as C
        from ._lib import lib
        self._lib, self._C = lib, C
        self.devices = list(devices)
        arr = (C.c_int * len(self.devices))(*self.devices)
        h = C.c_void_p()
        rc = lib.later_b200_mgpu_create(C.byref(h), len(self.devices), arr)
        if rc != 0:
            raise RuntimeError(f"later_b200_mgpu_create failed: {rc}")
        self._h = h

    def tsqr(self, m_local: int, n: int, A: list, lda: int, R: list, ldr: int) -> None:
        C = self._C
        P = len(self.devices)
        pa = (C.c_void_p * P)(*[a.data_ptr() for a in A])
        pr = (C.c_void_p * P)(*[r.data_ptr() for r in R])
        rc = self._lib.later_b200_tsqr_mgpu(self._h, m_local, n, pa, lda, pr, ldr)
        if rc != 0:
            raise RuntimeError(f"later_b200_tsqr_mgpu: {rc}: "
                               f"{self._lib.later_b200_mgpu_last_error(self._h).decode(errors='replace')}")

    def rgsqrf(self, m_local: int, n: int, A: list, lda: int, R: list, ldr: int) -> None:
        """The row-sharded recursion (later_b200_rgsqrf_mgpu): preferred - single-GPU accuracy, no
        redundant stack factorisation, no back-multiplication."""
        C = self._C
        P = len(self.devices)
        pa = (C.c_void_p * P)(*[a.data_ptr() for a in A])
        pr = (C.c_void_p * P)(*[r.data_ptr() for r in R])
        rc = self._lib.later_b200_rgsqrf_mgpu(self._h, m_local, n, pa, lda, pr, ldr)
        if rc != 0:
            raise RuntimeError(f"later_b200_rgsqrf_mgpu: {rc}: "
                               f"{self._lib.later_b200_mgpu_last_error(self._h).decode(errors='replace')}")

    def sync(self) -> None:
        rc = self._lib.later_b200_mgpu_sync(self._h)
        if rc != 0:
            raise RuntimeError(f"later_b200_mgpu_sync: {rc}")

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.later_b200_mgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
