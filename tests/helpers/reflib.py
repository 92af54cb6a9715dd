"""TEST INFRASTRUCTURE: ctypes binding of oracle/_ref/libref_later.so - the UNMODIFIED reference
compiled by oracle/Makefile - so that GPU parity tests can run the reference beside the product on the
same device buffer.  Only tests/ and bench.py's reference arm use it."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[2]
PATH = ROOT / "oracle" / "_ref" / "libref_later.so"


def available() -> bool:
    return PATH.exists()


class RefLib:
    def __init__(self):
        if not PATH.exists():
            raise FileNotFoundError(f"{PATH} not built (oracle/Makefile needs /root/reference)")
        self.lib = C.CDLL(str(PATH))
        vp, ci = C.c_void_p, C.c_int
        self.lib.ref_later_rgsqrf.argtypes = [ci, ci, vp, ci, vp, ci, vp, ci, vp, ci]
        self.lib.ref_mgs_caqr_panel_256x128.argtypes = [ci, ci, vp, ci, vp, ci, vp]
        for name in ("ref_mgs_caqr_panel_256x32",):
            if hasattr(self.lib, name):
                getattr(self.lib, name).argtypes = [ci, ci, vp, ci, vp, ci, vp]

    def houqr_q(self, A0: torch.Tensor, blocked: bool = False):
        """Reference later_rhouqr + later_ormqr (or later_bhouqr + later_ormqr2), as its driver chains them
        (test/test_qr.cu:116-127, :160-172); returns the explicit Q and R."""
        m, n = A0.shape
        vp, ci = C.c_void_p, C.c_int
        fn = self.lib.ref_later_bhouqr if blocked else self.lib.ref_later_rhouqr
        fn.argtypes = [ci, ci, vp, ci, vp, ci, vp, ci, vp, ci, vp, ci, vp]
        form = self.lib.ref_later_ormqr2 if blocked else self.lib.ref_later_ormqr
        form.argtypes = [ci, ci, vp, ci, vp, ci, vp]
        A = self._colmajor(A0)
        W = torch.zeros((n, m), device="cuda", dtype=torch.float32).t()
        R = torch.zeros((n, n), device="cuda", dtype=torch.float32).t()
        work = torch.zeros(m * n + (1 << 20), device="cuda")
        hwork = torch.zeros(m * n, device="cuda", dtype=torch.float16)
        U = torch.zeros(32 * 32, device="cuda")
        rc = fn(m, n, A.data_ptr(), m, W.data_ptr(), m, R.data_ptr(), n, work.data_ptr(), work.numel(),
                hwork.data_ptr(), hwork.numel(), U.data_ptr())
        torch.cuda.synchronize()
        if rc != 0:
            raise RuntimeError(f"reference householder qr: cuda error {rc}")
        rc = form(m, n, W.data_ptr(), m, A.data_ptr(), m, work.data_ptr())
        torch.cuda.synchronize()
        if rc != 0:
            raise RuntimeError(f"reference later_ormqr: cuda error {rc}")
        return W, torch.triu(R)

    def qdwh_polar(self, X0: torch.Tensor) -> torch.Tensor:
        """Reference later_qdwh_polar (EVD/later_qdwh_polar.cu:24) on a copy of X0 (n x n); returns U."""
        n = X0.shape[0]
        self.lib.ref_later_qdwh_polar.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        tmpA = self._colmajor(X0)
        A = torch.zeros((n, 2 * n), device="cuda", dtype=torch.float32).t()          # 2n x n, lda = 2n
        work = torch.zeros(max(n * n, 2 * n // 256 * 32 * n) + (1 << 20), device="cuda")
        hwork = torch.zeros(2 * n * n, device="cuda", dtype=torch.float16)
        rc = self.lib.ref_later_qdwh_polar(n, A.data_ptr(), 2 * n, tmpA.data_ptr(), work.data_ptr(), hwork.data_ptr())
        torch.cuda.synchronize()
        if rc != 0:
            raise RuntimeError(f"reference later_qdwh_polar: cuda error {rc}")
        return A[:n].clone()

    @staticmethod
    def _colmajor(x: torch.Tensor) -> torch.Tensor:
        out = torch.empty((x.shape[1], x.shape[0]), device="cuda", dtype=x.dtype).t()
        out.copy_(x)
        return out

    def rgsqrf(self, A0: torch.Tensor):
        """Reference later_rgsqrf on a copy of A0 (m x n, any layout); returns column-major Q, R."""
        m, n = A0.shape
        A = self._colmajor(A0)
        R = torch.zeros((n, n), device="cuda", dtype=torch.float32).t()
        work = torch.zeros(max(m // 256 * 32 * n, 1 << 20) + 4 * m + (1 << 20), device="cuda")
        hwork = torch.zeros(m * n, device="cuda", dtype=torch.float16)
        rc = self.lib.ref_later_rgsqrf(m, n, A.data_ptr(), m, R.data_ptr(), n, work.data_ptr(), work.numel(),
                                       hwork.data_ptr(), hwork.numel())
        torch.cuda.synchronize()
        if rc != 0:
            raise RuntimeError(f"reference later_rgsqrf: cuda error {rc}")
        return A, R

    def panel128(self, A0: torch.Tensor):
        m, n = A0.shape
        A = self._colmajor(A0)
        R = torch.zeros((n, n), device="cuda", dtype=torch.float32).t()
        work = torch.zeros(m * max(n, 32) + 65536, device="cuda")
        rc = self.lib.ref_mgs_caqr_panel_256x128(m, n, A.data_ptr(), m, R.data_ptr(), n, work.data_ptr())
        torch.cuda.synchronize()
        if rc != 0:
            raise RuntimeError(f"reference mgs_caqr_panel_256x128: cuda error {rc}")
        return A, R
