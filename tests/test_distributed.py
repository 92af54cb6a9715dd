"""The N > 1 paths on CPU, world_size 2, gloo backend: the host-side logic of the row-sharded TSQR
(later_b200/tsqr.py) with the numpy oracle standing in for the CUDA kernels through tsqr_rgsqrf's injection
points, and the algorithm of the row-sharded recursion (later_b200_rgsqrf_dist) as a numpy model with its
two kinds of all-reduce going through torch.distributed."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))


def _np_view(t):
    """numpy (m, n) view of a column-major torch CPU tensor (shares memory)."""
    m, n = t.shape
    ld = t.stride(1) if n > 1 else m
    base = t.t()  # (n, m) with strides (ld, 1)
    return np.lib.stride_tricks.as_strided(base.numpy(), shape=(m, n), strides=(4, 4 * ld))


def _worker(rank, world, port, m, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import rgsqrf_oracle as orc
    from later_b200.tsqr import tsqr_rgsqrf

    def oracle_qr(mm, nn, a, lda, r, ldr):
        av, rv = _np_view(a), _np_view(r)
        q, rr = orc.later_rgsqrf(av[:mm, :nn])
        av[:mm, :nn] = q
        rv[:nn, :nn] = rr

    def oracle_apply(mm, nn, q, ldq, w, ldw):
        qv, wv = _np_view(q), _np_view(w)
        qv[:mm, :nn] = (orc.s2h(qv[:mm, :nn]).astype(np.float32) @ orc.s2h(wv[:nn, :nn]).astype(np.float32))

    rng = np.random.default_rng(21)
    A_glob = rng.standard_normal((m, n), dtype=np.float32)
    mloc = m // world
    A_loc = torch.empty((n, mloc)).t()
    A_loc.copy_(torch.from_numpy(A_glob[rank * mloc:(rank + 1) * mloc]))
    R = torch.zeros((n, n)).t()
    tsqr_rgsqrf(mloc, n, A_loc, mloc, R, n, local_qr=oracle_qr, stack_qr=oracle_qr, apply_w=oracle_apply)
    np.save(os.path.join(tmp, f"q{rank}.npy"), np.array(_np_view(A_loc)))
    np.save(os.path.join(tmp, f"r{rank}.npy"), np.array(_np_view(R)))
    dist.barrier()
    dist.destroy_process_group()


def test_tsqr_two_ranks_gloo(tmp_path):
    from oracle import rgsqrf_oracle as orc
    m, n, world = 1024, 128, 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, m, n, str(tmp_path)), nprocs=world, join=True)
    Q = np.concatenate([np.load(tmp_path / f"q{r}.npy") for r in range(world)], axis=0)
    R0, R1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(R0, R1)                      # every rank holds the same R, bit for bit
    A = np.random.default_rng(21).standard_normal((m, n), dtype=np.float32)
    assert np.abs(np.tril(R0, -1)).max() == 0 and (np.diag(R0) > 0).all()
    assert orc.check_result(A, Q, R0) < 1e-3           # fp16 back-multiplication tolerance
    assert orc.check_otho(Q) < 1e-4
    # same factor as a single-process factorisation of the whole matrix, up to rounding
    _, Rref = orc.later_rgsqrf(A)
    assert np.abs(R0 - Rref).max() <= 5e-3 * np.abs(Rref).max()


def _sharded_worker(rank, world, port, m, n, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.helpers.sharded_model import rgsqrf_sharded
    calls = {"f64": 0, "f32": 0}

    def allreduce(x):
        calls["f64" if x.dtype == np.float64 else "f32"] += 1
        t = torch.from_numpy(np.ascontiguousarray(x))
        dist.all_reduce(t)
        return t.numpy()

    A_glob = np.random.default_rng(22).standard_normal((m, n), dtype=np.float32)
    mloc = m // world
    Q, R = rgsqrf_sharded(A_glob[rank * mloc:(rank + 1) * mloc], allreduce)
    np.save(os.path.join(tmp, f"q{rank}.npy"), Q)
    np.save(os.path.join(tmp, f"r{rank}.npy"), R)
    np.save(os.path.join(tmp, f"calls{rank}.npy"), np.array([calls["f64"], calls["f32"]]))
    dist.barrier()
    dist.destroy_process_group()


def test_row_sharded_recursion_two_ranks_gloo(tmp_path):
    """What later_b200_rgsqrf_dist promises: one fp64 all-reduce per panel and one fp32
    all-reduce per node, the same R on every rank bit for bit, and the accuracy of the unsharded factorisation
    (the TSQR variant rounds Q to fp16 once more and is 50x worse)."""
    from oracle import rgsqrf_oracle as orc
    from tests.helpers.sharded_model import rgsqrf_sharded
    m, n, world = 2048, 512, 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_sharded_worker, args=(world, port, m, n, str(tmp_path)), nprocs=world, join=True)
    Q = np.concatenate([np.load(tmp_path / f"q{r}.npy") for r in range(world)], axis=0)
    R0, R1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(R0, R1)
    assert np.load(tmp_path / "calls0.npy").tolist() == [n // 128, n // 128 - 1]     # panels, nodes
    A = np.random.default_rng(22).standard_normal((m, n), dtype=np.float32)
    assert np.abs(np.tril(R0, -1)).max() == 0 and (np.diag(R0) > 0).all()
    Q1, R1p = rgsqrf_sharded(A)                        # one rank: no exchange at all
    back1, orth1 = orc.check_result(A, Q1, R1p), orc.check_otho(Q1)
    back, orth = orc.check_result(A, Q, R0), orc.check_otho(Q)
    assert back <= 1.1 * back1 and orth <= 1.1 * orth1
    assert np.abs(R0 - R1p).max() <= 1e-3 * np.abs(R1p).max()
    # ... which is the accuracy of the reference's algorithm on the same matrix (oracle, fp32 MGS panels)
    Qo, Ro = orc.later_rgsqrf(A)
    assert back <= 2 * orc.check_result(A, Qo, Ro) and orth <= 2 * orc.check_otho(Qo)


def test_stack_layout():
    from later_b200.tsqr import stack_from_gathered
    P, n = 3, 4
    Rs = [torch.arange(n * n, dtype=torch.float32).reshape(n, n) + 100 * p for p in range(P)]  # R_p[i, j]
    gathered = torch.stack([r.t().contiguous() for r in Rs])      # storage of column-major R_p
    S = stack_from_gathered(gathered)
    assert S.shape == (P * n, n) and S.stride() == (1, P * n)
    for p in range(P):
        assert torch.equal(S[p * n:(p + 1) * n, :], Rs[p])


def test_host_A_requires_builtin_local_qr():
    """`host_A` streams the row block through later_b200_rgsqrf_stream_in; it cannot be combined with
    an injected local_qr (the CPU stand-ins of these tests)."""
    import pytest, torch
    from later_b200.tsqr import tsqr_rgsqrf
    A = torch.zeros(256, 128).t().contiguous().t()
    R = torch.zeros(128, 128)
    with pytest.raises(ValueError, match="host_A needs the built-in"):
        tsqr_rgsqrf(256, 128, A, 256, R, 128, local_qr=lambda *a: None, stack_qr=lambda *a: None,
                    apply_w=lambda *a: None, host_A=A)
