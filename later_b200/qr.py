"""Host-side mirror of the reference's RGSQRF interface (reference include/LATER.h:39-47) on top of
the C ABI.  Names, argument order and meaning follow the reference; matrices are column-major fp32
device buffers (torch tensors of shape (m, n) with strides (1, lda)).

PyTorch supplies device memory and streams only; all arithmetic runs in liblater_b200.so.
"""
from __future__ import annotations

import ctypes as C

import torch

from ._lib import lib


class LaterError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"later_b200 error {code}: {message}")
        self.code = code


def colmajor_empty(m: int, n: int, device="cuda", dtype=torch.float32, ld: int | None = None):
    """An (m, n) tensor stored column-major with leading dimension ld (default m)."""
    ld = m if ld is None else ld
    if ld < m:
        raise ValueError("ld < m")
    return torch.empty((n, ld), device=device, dtype=dtype).t()[:m, :]


def to_colmajor(x: torch.Tensor) -> torch.Tensor:
    """Copy of a 2-D tensor in column-major storage."""
    out = colmajor_empty(x.shape[0], x.shape[1], device=x.device, dtype=x.dtype)
    out.copy_(x)
    return out


def _check_colmajor(name: str, t: torch.Tensor, rows: int, cols: int, ld: int, dtype=torch.float32):
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}")
    if t.dim() != 2 or t.shape[0] < rows or t.shape[1] < cols:
        raise ValueError(f"{name} has shape {tuple(t.shape)}, need at least ({rows}, {cols})")
    if cols > 1 and (t.stride(0) != 1 or t.stride(1) != ld):
        raise ValueError(f"{name} must be column-major with leading dimension {ld}; "
                         f"strides are {t.stride()}")


class Context:
    """One device + one stream (reference: struct cudaCtxt, include/LATER.h:19-22, minus the
    cuBLAS/cuSOLVER handles, which this implementation does not need)."""

    def __init__(self, device: int | None = None, stream: torch.cuda.Stream | None = None,
                 use_graph: bool = True):
        if not torch.cuda.is_available():
            raise LaterError(-4, "no CUDA device: later_b200 has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        with torch.cuda.device(self.device):
            s = torch.cuda.current_stream() if stream is None else stream
            self._stream = s
            h = C.c_void_p()
            rc = lib.later_b200_create(C.byref(h), self.device, C.c_void_p(s.cuda_stream))
        if rc != 0:
            raise LaterError(rc, "later_b200_create failed (an sm_100 GPU is required)")
        self._h = h
        if not use_graph:
            lib.later_b200_set_graph(self._h, 0)

    def close(self):
        if getattr(self, "_h", None):
            lib.later_b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise(self, rc: int):
        raise LaterError(rc, lib.later_b200_last_error(self._h).decode(errors="replace"))

    @property
    def last_launch_count(self) -> int:
        return int(lib.later_b200_last_launch_count(self._h))

    def workspace_bytes(self, m: int, n: int) -> int:
        return int(lib.later_b200_workspace_bytes(self._h, m, n))

    def last_info(self) -> dict:
        """Numerical status of the most recent factorisation (waits for the stream); see
        later_b200_last_info in include/later_b200.h."""
        arr = (C.c_int * 4)()
        rc = lib.later_b200_last_info(self._h, arr)
        if rc not in (0, -5):
            self._raise(rc)
        return {"status": rc, "bad_column": arr[0] - 1 if arr[0] else None, "nonfinite": bool(arr[1] & 1),
                "fallback_panels": arr[2], "cond_log2": arr[3]}

    def graph_stats(self) -> tuple[int, int]:
        r, c = C.c_long(), C.c_long()
        lib.later_b200_graph_stats(self._h, C.byref(r), C.byref(c))
        return int(r.value), int(c.value)


_default: dict[int, Context] = {}


def default_context() -> Context:
    dev = torch.cuda.current_device()
    if dev not in _default:
        _default[dev] = Context(dev)
    return _default[dev]


def later_rgsqrf(ctxt: Context | None, m: int, n: int, A: torch.Tensor, lda: int, R: torch.Tensor,
                 ldr: int, work=None, lwork: int = 0, hwork=None, lhwork: int = 0) -> None:
    """A = Q R.  A (m x n, column-major, lda) is overwritten by the explicit Q; R (n x n, ldr)
    receives the upper-triangular factor.  work/hwork are accepted and ignored, as in the C++
    wrapper.  Reference: later_rgsqrf, QR/later_rgsqrf.cu:62-79."""
    ctxt = ctxt or default_context()
    _check_colmajor("A", A, m, n, lda)
    _check_colmajor("R", R, n, n, ldr)
    rc = lib.later_b200_rgsqrf(ctxt._h, m, n, A.data_ptr(), lda, R.data_ptr(), ldr)
    if rc != 0:
        ctxt._raise(rc)


def later_rgsqrf_reorth(ctxt: Context | None, m: int, n: int, A: torch.Tensor, lda: int, R: torch.Tensor,
                        ldr: int) -> None:
    """Gram-Schmidt twice: A = Q1 R1, Q1 = Q2 R2; A <- Q2, R <- R2 R1 (include/later_b200.h)."""
    ctxt = ctxt or default_context()
    _check_colmajor("A", A, m, n, lda)
    _check_colmajor("R", R, n, n, ldr)
    rc = lib.later_b200_rgsqrf_reorth(ctxt._h, m, n, A.data_ptr(), lda, R.data_ptr(), ldr)
    if rc != 0:
        ctxt._raise(rc)


def later_qdwh_polar(ctxt: Context | None, n: int, A: torch.Tensor, lda: int, H, ldh: int, tmpA: torch.Tensor,
                     work=None, hwork=None, smin_est: float = 0.0, max_iter: int = 0) -> int:
    """Polar factor by the QDWH iteration (reference EVD/later_qdwh_polar.cu:24-110), with the
    reference's argument list: tmpA (n x n, ld n) holds the matrix (normalised in place), A (2n x n,
    lda) is the stacked workspace whose top block receives U; H, work and hwork are accepted and
    ignored (the reference never writes H either).  Returns the number of iterations."""
    ctxt = ctxt or default_context()
    _check_colmajor("tmpA", tmpA, n, n, tmpA.stride(1))
    _check_colmajor("A", A, 2 * n, n, lda)
    it = C.c_int(0)
    rc = lib.later_b200_qdwh_polar(ctxt._h, n, tmpA.data_ptr(), tmpA.stride(1), A.data_ptr(), lda,
                                   float(smin_est), int(max_iter), C.byref(it))
    if rc != 0:
        ctxt._raise(rc)
    return int(it.value)


def comm_init(ctxt: Context, group=None, peer_message_bytes: int = 4 << 20) -> None:
    """Gives the context an NCCL communicator over the ranks of a torch.distributed group (one rank per
    GPU): rank 0 draws the NCCL unique id, torch.distributed carries its 128 bytes to the others."""
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        rc = lib.later_b200_comm_unique_id(buf)
        if rc != 0:
            raise LaterError(rc, "later_b200_comm_unique_id failed (libnccl.so.2 not loadable?)")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=f"cuda:{ctxt.device}")
    dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    raw = bytes(t.cpu().tolist())
    rc = lib.later_b200_comm_init(ctxt._h, world, rank, raw)
    if rc != 0:
        ctxt._raise(rc)
    # NVLink peer-memory path for the (small, latency-bound) all-reduces: exchange the slabs' IPC handles
    hbuf = (C.c_ubyte * 64)()
    rc = lib.later_b200_peer_export(ctxt._h, peer_message_bytes, hbuf)
    if rc != 0:
        ctxt._raise(rc)
    mine = torch.tensor(list(hbuf), dtype=torch.uint8, device=f"cuda:{ctxt.device}")
    allh = torch.empty(world * 64, dtype=torch.uint8, device=f"cuda:{ctxt.device}")
    dist.all_gather_into_tensor(allh, mine, group=group)
    rc = lib.later_b200_peer_import(ctxt._h, world, rank, bytes(allh.cpu().tolist()))
    if rc != 0:
        ctxt._raise(rc)
    dist.barrier(group=group)


def later_rgsqrf_dist(ctxt: Context, m_local: int, n: int, A: torch.Tensor, lda: int, R: torch.Tensor,
                      ldr: int) -> None:
    """Row-sharded RGSQRF (collective over the ranks of comm_init): A (this rank's m_local x n row block)
    <- its rows of the global Q, R <- the global R, identical on every rank (include/later_b200.h)."""
    _check_colmajor("A", A, m_local, n, lda)
    _check_colmajor("R", R, n, n, ldr)
    rc = lib.later_b200_rgsqrf_dist(ctxt._h, m_local, n, A.data_ptr(), lda, R.data_ptr(), ldr)
    if rc != 0:
        ctxt._raise(rc)


def later_rgsqrf_host(ctxt: Context | None, m: int, n: int, A: torch.Tensor, lda: int,
                      R: torch.Tensor, ldr: int) -> None:
    """Same with HOST (ideally pinned) buffers; H2D and D2H copies happen inside the call, overlapped
    with the factorisation.  R's blocks below the block diagonal (granularity max(min(n, 256), n/16)) are
    not written (include/later_b200.h)."""
    ctxt = ctxt or default_context()
    if A.is_cuda or R.is_cuda:
        raise ValueError("later_rgsqrf_host takes host tensors")
    _check_colmajor("A", A, m, n, lda)
    _check_colmajor("R", R, n, n, ldr)
    rc = lib.later_b200_rgsqrf_host(ctxt._h, m, n, A.data_ptr(), lda, R.data_ptr(), ldr)
    if rc != 0:
        ctxt._raise(rc)


def later_oc_qr(ctxt: Context | None, m: int, n: int, A: torch.Tensor, lda: int, R: torch.Tensor, ldr: int,
                block_cols: int = 8192) -> None:
    """Out-of-core QR of a HOST matrix (reference later_oc_qr_rec / _blk, QR/later_oc_qr.cu:29-121):
    column blocks of block_cols stream through the device; A <- Q, R <- block upper triangle."""
    ctxt = ctxt or default_context()
    if A.is_cuda or R.is_cuda:
        raise ValueError("later_oc_qr takes host tensors")
    _check_colmajor("A", A, m, n, lda)
    _check_colmajor("R", R, n, n, ldr)
    rc = lib.later_b200_oc_qr(ctxt._h, m, n, A.data_ptr(), lda, R.data_ptr(), ldr, block_cols)
    if rc != 0:
        ctxt._raise(rc)


def later_rgsqrf_stream_in(ctxt: Context | None, m: int, n: int, hA: torch.Tensor, hlda: int,
                           A: torch.Tensor, lda: int, R: torch.Tensor, ldr: int) -> None:
    """Host in, device out: hA (host, ideally pinned) is copied into A (device) column piece by
    column piece and factored as it arrives; Q (in A) and R stay on the device.  Asynchronous on the
    context's stream; hA must stay alive until the stream has passed the call."""
    ctxt = ctxt or default_context()
    if hA.is_cuda or not A.is_cuda or not R.is_cuda:
        raise ValueError("later_rgsqrf_stream_in takes a host hA and device A, R")
    _check_colmajor("hA", hA, m, n, hlda)
    _check_colmajor("A", A, m, n, lda)
    _check_colmajor("R", R, n, n, ldr)
    rc = lib.later_b200_rgsqrf_stream_in(ctxt._h, m, n, hA.data_ptr(), hlda, A.data_ptr(), lda,
                                         R.data_ptr(), ldr)
    if rc != 0:
        ctxt._raise(rc)


def mgs_caqr_panel_256x128(ctxt: Context | None, m: int, n: int, A: torch.Tensor, lda: int,
                           R: torch.Tensor, ldr: int, work=None) -> None:
    """QR of an m x 128 panel (reference QR/panel.cu:10-63)."""
    ctxt = ctxt or default_context()
    _check_colmajor("A", A, m, n, lda)
    _check_colmajor("R", R, n, n, ldr)
    rc = lib.later_b200_panel_qr(ctxt._h, m, n, A.data_ptr(), lda, R.data_ptr(), ldr)
    if rc != 0:
        ctxt._raise(rc)


def mgs_caqr_panel_256x32(ctxt: Context | None, m: int, n: int, A: torch.Tensor, lda: int,
                          R: torch.Tensor, ldr: int, work=None) -> None:
    """QR of an m x 32 strip (reference QR/panel.cu:65-134)."""
    ctxt = ctxt or default_context()
    _check_colmajor("A", A, m, n, lda)
    _check_colmajor("R", R, min(m, n), min(m, n), ldr)
    rc = lib.later_b200_panel32_qr(ctxt._h, m, n, A.data_ptr(), lda, R.data_ptr(), ldr)
    if rc != 0:
        ctxt._raise(rc)


def later_rhouqr(ctxt: Context | None, m: int, n: int, A: torch.Tensor, lda: int, W: torch.Tensor, ldw: int,
                 R: torch.Tensor, ldr: int, work=None, lwork: int = 0, hwork=None, lhwork: int = 0, U=None,
                 merge_top: bool = False) -> None:
    """Recursive Householder QR in WY form (reference QR/later_rhouqr.cu:21-201): A <- Y, W, R out;
    later_ormqr(m, n, W, ldw, A, lda) then forms the explicit Q.  merge_top=True is later_bhouqr's
    contract (complete W, for later_ormqr2)."""
    ctxt = ctxt or default_context()
    _check_colmajor("A", A, m, n, lda)
    _check_colmajor("W", W, m, n, ldw)
    _check_colmajor("R", R, n, n, ldr)
    rc = lib.later_b200_rhouqr(ctxt._h, m, n, A.data_ptr(), lda, W.data_ptr(), ldw, R.data_ptr(), ldr,
                               1 if merge_top else 0)
    if rc != 0:
        ctxt._raise(rc)


def later_ormqr(m: int, n: int, W: torch.Tensor, ldw: int, Y: torch.Tensor, ldy: int, work=None,
                ctxt: Context | None = None) -> None:
    """W <- explicit Q from the WY pair (reference QR/later_ormqr.cu:18-64)."""
    ctxt = ctxt or default_context()
    _check_colmajor("W", W, m, n, ldw)
    _check_colmajor("Y", Y, m, n, ldy)
    rc = lib.later_b200_ormqr(ctxt._h, m, n, W.data_ptr(), ldw, Y.data_ptr(), ldy)
    if rc != 0:
        ctxt._raise(rc)


def later_ormqr2(m: int, n: int, W: torch.Tensor, ldw: int, Y: torch.Tensor, ldy: int, work=None,
                 ctxt: Context | None = None) -> None:
    """W <- I - W Y(0:n,0:n)^T only (reference QR/later_ormqr.cu:66-85)."""
    ctxt = ctxt or default_context()
    _check_colmajor("W", W, m, n, ldw)
    _check_colmajor("Y", Y, m, n, ldy)
    rc = lib.later_b200_ormqr2(ctxt._h, m, n, W.data_ptr(), ldw, Y.data_ptr(), ldy)
    if rc != 0:
        ctxt._raise(rc)


def tsqr_apply(ctxt: Context, m: int, n: int, Q: torch.Tensor, ldq: int, W: torch.Tensor,
               ldw: int) -> None:
    _check_colmajor("Q", Q, m, n, ldq)
    _check_colmajor("W", W, n, n, ldw)
    rc = lib.later_b200_tsqr_apply(ctxt._h, m, n, Q.data_ptr(), ldq, W.data_ptr(), ldw)
    if rc != 0:
        ctxt._raise(rc)


# ---- diagnostics -------------------------------------------------------------------------------
def gemm_gram(ctxt: Context, Qh: torch.Tensor, colA: int, Mc: int, colB: int, Nc: int,
              C_out: torch.Tensor, Ch_out: torch.Tensor | None = None, splits: int = 0) -> None:
    """C = Qh[:, colA:colA+Mc]^T Qh[:, colB:colB+Nc] (fp16 operands, fp32 accumulate)."""
    _check_colmajor("Qh", Qh, Qh.shape[0], Qh.shape[1], Qh.stride(1), torch.float16)
    _check_colmajor("C", C_out, Mc, Nc, C_out.stride(1))
    rc = lib.later_b200_gemm_gram(
        ctxt._h, Qh.data_ptr(), Qh.shape[0], Qh.shape[1], Qh.stride(1), colA, Mc, colB, Nc,
        C_out.data_ptr(), C_out.stride(1),
        Ch_out.data_ptr() if Ch_out is not None else None,
        Ch_out.stride(1) if Ch_out is not None else 0, splits)
    if rc != 0:
        ctxt._raise(rc)


def gemm_update(ctxt: Context, Qh: torch.Tensor, colA: int, K: int, Bh: torch.Tensor,
                C_io: torch.Tensor, Ch_out: torch.Tensor | None = None, subtract: bool = True) -> None:
    """C (-)= Qh[:, colA:colA+K] Bh   (Bh: K x Nc fp16 column-major)."""
    Nc = Bh.shape[1]
    _check_colmajor("Qh", Qh, Qh.shape[0], Qh.shape[1], Qh.stride(1), torch.float16)
    _check_colmajor("Bh", Bh, K, Nc, Bh.stride(1), torch.float16)
    _check_colmajor("C", C_io, Qh.shape[0], Nc, C_io.stride(1))
    rc = lib.later_b200_gemm_update(
        ctxt._h, Qh.data_ptr(), Qh.shape[0], Qh.shape[1], Qh.stride(1), colA, K, Bh.data_ptr(),
        Bh.stride(1), Nc, C_io.data_ptr(), C_io.stride(1),
        Ch_out.data_ptr() if Ch_out is not None else None,
        Ch_out.stride(1) if Ch_out is not None else 0, 1 if subtract else 0)
    if rc != 0:
        ctxt._raise(rc)


# ---- the reference driver's self-consistency metrics (reference test/test_qr.cu:216-268) --------
def backward_error(A0: torch.Tensor, Q: torch.Tensor, R: torch.Tensor, dtype=torch.float32) -> float:
    """||A - Q R||_F / ||A||_F (checkResult, reference test/test_qr.cu:216-228)."""
    res = A0.to(dtype) - Q.to(dtype) @ R.to(dtype)
    return float(torch.linalg.norm(res.double()) / torch.linalg.norm(A0.double()))


def orthogonality(Q: torch.Tensor, dtype=torch.float32) -> float:
    """||I - Q^T Q||_F / n - note the division by n (checkOtho, reference test/test_qr.cu:245-268)."""
    n = Q.shape[1]
    G = Q.to(dtype).t() @ Q.to(dtype)
    G.diagonal().sub_(1.0)
    return float(torch.linalg.norm(G.double()) / n)
