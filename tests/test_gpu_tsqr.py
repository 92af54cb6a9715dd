"""Multi-GPU TSQR on real hardware: 2 ranks over NCCL (skipped on a single-GPU box)."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n", [(16384, 256), (65536, 1024)])
def test_tsqr_two_gpus(m, n):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29611",
           str(ROOT / "tests" / "helpers" / "tsqr_worker.py"), str(m), str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout


def test_tsqr_mgpu_c_abi_from_a_cxx_caller(tmp_path):
    """later_b200_tsqr_mgpu (single process, one host thread, P devices, NCCL all-gather) driven by a
    C++ program that includes only include/later_b200.h.  P = 1 runs everywhere; P = 2 needs 2 GPUs."""
    exe = tmp_path / "test_tsqr_mgpu"
    r = subprocess.run(["nvcc", "-std=c++17", "-O2", "-I", str(ROOT / "include"), str(ROOT / "tests/c/test_tsqr_mgpu.cu"),
                        str(ROOT / "later_b200/liblater_b200.so"), "-Xlinker", f"-rpath={ROOT / 'later_b200'}",
                        "-o", str(exe)], capture_output=True, text=True, timeout=180)
    assert r.returncode == 0, r.stderr[-2000:]
    for P in ([1, 2] if torch.cuda.device_count() >= 2 else [1]):
        for variant in ([], ["tsqr"]):          # later_b200_rgsqrf_mgpu (default), later_b200_tsqr_mgpu
            r = subprocess.run([str(exe), str(P), "16384", "256"] + variant, capture_output=True, text=True, timeout=180)
            assert r.returncode == 0 and "OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_tsqr_mgpu_matches_the_single_gpu_factorisation():
    """The same entry point through the Python binding: R equals the single-GPU factor of the whole
    matrix up to rounding, every device holds the same bits."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    sys.path.insert(0, str(ROOT))
    from later_b200 import qr
    from later_b200.tsqr import MultiGpu
    P, m_loc, n = 2, 65536, 512
    g = torch.Generator(device="cuda:0").manual_seed(5)
    A_glob = torch.randn(P * m_loc, n, device="cuda:0", generator=g)
    A, R = [], []
    for p in range(P):
        with torch.cuda.device(p):
            a = torch.empty((n, m_loc), device=f"cuda:{p}").t()
            a.copy_(A_glob[p * m_loc:(p + 1) * m_loc].to(f"cuda:{p}"))
            A.append(a)
            R.append(torch.empty((n, n), device=f"cuda:{p}").t())
    with torch.cuda.device(0):
        c = qr.Context(0)
        A1 = qr.to_colmajor(A_glob)
        R1 = qr.colmajor_empty(n, n)
        qr.later_rgsqrf(c, P * m_loc, n, A1, P * m_loc, R1, n)
        back1, orth1 = qr.backward_error(A_glob, A1, R1), qr.orthogonality(A1)
        c.close()
    mg = MultiGpu(list(range(P)))
    for variant in ("rgsqrf", "tsqr"):
        for p in range(P):
            A[p].copy_(A_glob[p * m_loc:(p + 1) * m_loc].to(f"cuda:{p}"))
        getattr(mg, variant)(m_loc, n, A, m_loc, R, n)
        mg.sync()
        assert all(torch.equal(R[0].cpu(), r.cpu()) for r in R[1:])
        Q = torch.cat([a.to("cuda:0") for a in A], dim=0)
        back, orth = qr.backward_error(A_glob, Q, R[0]), qr.orthogonality(Q)
        if variant == "rgsqrf":     # the row-sharded recursion is as accurate as one GPU
            assert back <= 2 * back1 + 1e-7 and orth <= 2 * orth1 + 1e-8, (back, back1, orth, orth1)
        else:                       # the TSQR variant rounds Q and W to fp16 once more
            assert back <= 5e-4 and orth <= 5e-5
        assert (R[0] - R1).abs().max().item() <= 5e-3 * R1.abs().max().item()
    mg.close()
