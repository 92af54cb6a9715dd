"""Parity tests proper: the CUDA path (through the C ABI) against the oracle on identical seeded
inputs, against the committed outputs of the reference itself, and - at BASELINE.json's full sizes,
where the oracle would take too long - through size-independent properties of a QR factorisation.

Tolerances (floating point, so not bit-exact; BASELINE.json's bar):
  * backward error ||A-QR||/||A|| and orthogonality ||I-Q^T Q||/n each within 2x of the
    reference's (oracle or golden) value on the same input;
  * fp16-tensor-core tolerance vs LAPACK: backward error <= 1e-3 (~2 u_fp16).
"""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import rgsqrf_oracle as orc  # noqa: E402
from tests.golden.make_golden import CASES, make_input  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = ROOT / "tests" / "golden"
META = json.loads((GOLD / "golden_meta.json").read_text())


@pytest.fixture(scope="module")
def qr():
    from later_b200 import qr as _qr
    torch.backends.cuda.matmul.allow_tf32 = False
    return _qr


@pytest.fixture(scope="module")
def ctx(qr):
    c = qr.Context()
    yield c
    c.close()


def dev_colmajor(qr, a: np.ndarray) -> torch.Tensor:
    t = qr.colmajor_empty(a.shape[0], a.shape[1])
    t.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))
    return t


def run_rgsqrf(qr, ctx, A0: np.ndarray):
    m, n = A0.shape
    A = dev_colmajor(qr, A0)
    R = qr.colmajor_empty(n, n)
    R.fill_(float("nan"))                     # every entry of R must be written by the library
    qr.later_rgsqrf(ctx, m, n, A, m, R, n)
    torch.cuda.synchronize()
    return A.cpu().numpy(), R.cpu().numpy()


# ------------------------------------------------------------------------------ GEMM kernels
@pytest.mark.parametrize("m,Mc,Nc,splits", [(1024, 128, 128, 1), (1024, 256, 256, 1),
                                            (4096, 128, 128, 8), (8200, 512, 512, 0),
                                            (2048, 1024, 1024, 1), (65536, 128, 128, 0)])
def test_gram_kernel_vs_fp32_reference(qr, ctx, m, Mc, Nc, splits):
    g = torch.Generator(device="cuda").manual_seed(1)
    Q = qr.to_colmajor(torch.randn(m, Mc + Nc, device="cuda", generator=g).half())
    C = qr.colmajor_empty(Mc, Nc)
    Ch = qr.colmajor_empty(Mc, Nc, dtype=torch.float16)
    C.fill_(float("nan"))
    qr.gemm_gram(ctx, Q, 0, Mc, Mc, Nc, C, Ch, splits)
    ref = Q[:, :Mc].float().t() @ Q[:, Mc:].float()          # plain fp32 reference, same fp16 inputs
    scale = ref.abs().max().item()
    assert (C - ref).abs().max().item() <= 2e-5 * scale      # fp32 accumulation-order noise only
    assert torch.equal(Ch, C.half())                          # fp16 copy is the RN cast of the result


@pytest.mark.parametrize("m,K,Nc,sub", [(1024, 128, 128, True), (1000, 256, 256, True),
                                        (4096, 512, 512, True), (2048, 256, 256, False),
                                        (16384, 1024, 128, True)])
def test_update_kernel_vs_fp32_reference(qr, ctx, m, K, Nc, sub):
    g = torch.Generator(device="cuda").manual_seed(2)
    Q = qr.to_colmajor(torch.randn(m, K + 64, device="cuda", generator=g).half())
    B = qr.to_colmajor((torch.randn(K, Nc, device="cuda", generator=g) / 8).half())
    C0 = qr.to_colmajor(torch.randn(m, Nc, device="cuda", generator=g))
    C = qr.to_colmajor(C0.clone())
    Ch = qr.colmajor_empty(m, Nc, dtype=torch.float16)
    qr.gemm_update(ctx, Q, 64, K, B, C, Ch, sub)
    prod = Q[:, 64:64 + K].float() @ B.float()
    ref = C0 - prod if sub else prod
    assert (C - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    assert torch.equal(Ch, C.half())


# ------------------------------------------------------------------------------ panel
@pytest.mark.parametrize("m,dist", [(128, "normal"), (256, "uniform"), (320, "normal"),
                                    (1024, "uniform"), (5000, "normal"), (65536, "uniform")])
def test_panel_vs_oracle(qr, ctx, m, dist):
    rng = np.random.default_rng(m)
    A0 = rng.random((m, 128), dtype=np.float32) if dist == "uniform" else rng.standard_normal((m, 128), dtype=np.float32)
    A = dev_colmajor(qr, A0)
    R = qr.colmajor_empty(128, 128)
    R.fill_(float("nan"))
    qr.mgs_caqr_panel_256x128(ctx, m, 128, A, m, R, 128)
    Q, R = A.cpu().numpy(), R.cpu().numpy()
    assert np.abs(np.tril(R, -1)).max() == 0.0 and (np.diag(R) > 0).all()
    Qo = np.array(A0, order="F", copy=True)
    Ro = np.zeros((128, 128), dtype=np.float32)
    orc.mgs_caqr_panel_256x128(Qo, Ro)
    # same fp32-level quality as the reference's MGS/CAQR panel ...
    assert orc.check_result(A0, Q, R) <= max(2 * orc.check_result(A0, Qo, Ro), 5e-7)
    assert orc.check_otho(Q) <= max(2 * orc.check_otho(Qo), 5e-8)
    # ... and the same factors (QR with r_ii > 0 is unique), up to fp32 rounding times conditioning
    cond = np.linalg.cond(A0.astype(np.float64))
    assert np.abs(R - Ro).max() <= 1e-5 * cond * np.abs(Ro).max()
    assert np.abs(Q - Qo).max() <= 1e-5 * cond


# ------------------------------------------------------------------------------ RGSQRF vs oracle
@pytest.mark.parametrize("m,n,dist", [(256, 256, "normal"), (512, 256, "normal"), (1024, 512, "uniform"),
                                      (1024, 1024, "uniform"), (2048, 1024, "normal"), (800, 128, "normal"),
                                      (1000, 256, "normal")])
def test_rgsqrf_vs_oracle(qr, ctx, m, n, dist):
    rng = np.random.default_rng(1000 + m + n)
    A0 = rng.random((m, n), dtype=np.float32) if dist == "uniform" else rng.standard_normal((m, n), dtype=np.float32)
    Q, R = run_rgsqrf(qr, ctx, A0)
    Qo, Ro = orc.later_rgsqrf(A0)
    assert np.isfinite(R).all() and np.abs(np.tril(R, -1)).max() == 0.0 and (np.diag(R) > 0).all()
    back, orth = orc.check_result(A0, Q, R), orc.check_otho(Q)
    back_o, orth_o = orc.check_result(A0, Qo, Ro), orc.check_otho(Qo)
    assert back <= 2 * back_o + 1e-7, (back, back_o)
    assert orth <= 2 * orth_o + 1e-7, (orth, orth_o)
    assert back <= 1e-3                                       # fp16-TC tolerance vs LAPACK (~5e-7)
    cond = np.linalg.cond(A0.astype(np.float64))
    assert np.abs(R - Ro).max() <= 2e-3 * cond * np.abs(Ro).max()


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v[0] == "rgsqrf"])
def test_rgsqrf_vs_reference_golden(qr, ctx, name):
    kind, m, n, dist, seed = CASES[name]
    g = np.load(GOLD / f"{name}.npz")
    A0 = make_input(kind, m, n, dist, seed)
    Q, R = run_rgsqrf(qr, ctx, A0)
    assert orc.check_result(A0, Q, R) <= 2 * float(g["backward"])
    assert orc.check_otho(Q) <= 2 * float(g["orth"])
    assert np.abs(np.triu(R) - np.triu(g["R"])).max() <= 2e-3 * np.abs(g["R"]).max()
    assert np.abs(Q[::8, :] - g["Q_rows8"]).max() <= 2e-3


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v[0] == "panel"])
def test_panel_vs_reference_golden(qr, ctx, name):
    kind, m, n, dist, seed = CASES[name]
    g = np.load(GOLD / f"{name}.npz")
    A0 = make_input(kind, m, n, dist, seed)
    A = dev_colmajor(qr, A0)
    R = qr.colmajor_empty(128, 128)
    qr.mgs_caqr_panel_256x128(ctx, m, 128, A, m, R, 128)
    Q, R = A.cpu().numpy(), R.cpu().numpy()
    assert orc.check_result(A0, Q, R) <= max(2 * float(g["backward"]), 5e-7)
    assert orc.check_otho(Q) <= max(2 * float(g["orth"]), 5e-8)
    # same factors up to fp32 rounding times conditioning (the sweep cases carry their cond)
    tol = 2e-5 * max(1.0, META.get(name, {}).get("cond", 1.0) / 10.0)
    if tol < 0.1:
        assert np.abs(np.triu(R) - np.triu(g["R"])).max() <= tol * np.abs(g["R"]).max()
        assert np.abs(Q[::8, :] - g["Q_rows8"]).max() <= tol


# ------------------------------------------------------------------------------ 32-column entry points
@pytest.mark.parametrize("m,dist", [(32, "normal"), (100, "normal"), (256, "uniform"), (257, "normal"),
                                    (1000, "normal"), (4096, "uniform"), (70000, "normal")])
def test_panel32_vs_oracle(qr, ctx, m, dist):
    """mgs_caqr_panel_256x32 (reference QR/panel.cu:65-134): same factors as the reference's CAQR tree
    (QR with r_ii > 0 is unique) at fp32 level, including single-block, ragged and one-row-over shapes."""
    rng = np.random.default_rng(500 + m)
    A0 = rng.random((m, 32), dtype=np.float32) if dist == "uniform" else rng.standard_normal((m, 32), dtype=np.float32)
    A = dev_colmajor(qr, A0)
    R = qr.colmajor_empty(32, 32)
    R.fill_(float("nan"))
    qr.mgs_caqr_panel_256x32(ctx, m, 32, A, m, R, 32)
    Q, R = A.cpu().numpy(), R.cpu().numpy()
    assert np.abs(np.tril(R, -1)).max() == 0.0 and (np.diag(R) > 0).all()
    Qo = np.array(A0, order="F", copy=True)
    Ro = orc.mgs_caqr_panel_256x32(Qo)
    assert orc.check_result(A0, Q, R) <= max(2 * orc.check_result(A0, Qo, Ro), 5e-7)
    assert orc.check_otho(Q) <= max(2 * orc.check_otho(Qo), 5e-8)
    cond = np.linalg.cond(A0.astype(np.float64))
    assert np.abs(R - Ro).max() <= 1e-5 * cond * np.abs(Ro).max()
    assert np.abs(Q - Qo).max() <= 1e-5 * cond


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v[0] == "panel32"])
def test_panel32_vs_reference_golden(qr, ctx, name):
    kind, m, n, dist, seed = CASES[name]
    g = np.load(GOLD / f"{name}.npz")
    A0 = make_input(kind, m, n, dist, seed)
    A = dev_colmajor(qr, A0)
    R = qr.colmajor_empty(32, 32)
    qr.mgs_caqr_panel_256x32(ctx, m, 32, A, m, R, 32)
    Q, R = A.cpu().numpy(), R.cpu().numpy()
    assert orc.check_result(A0, Q, R) <= max(2 * float(g["backward"]), 5e-7)
    assert orc.check_otho(Q) <= max(2 * float(g["orth"]), 5e-8)
    assert np.abs(np.triu(R) - np.triu(g["R"])).max() <= 2e-5 * np.abs(g["R"]).max()
    assert np.abs(Q[::8, :] - g["Q_rows8"]).max() <= 2e-5


def _read_csv(path):
    return np.loadtxt(path, delimiter=",", ndmin=2)


def test_reference_panel_drivers_run_unchanged_against_this_library(tmp_path):
    """test/test_mgs_panel.cu and test/test_caqr_panel.cu of the reference, compiled unmodified against
    later_b200 (oracle/Makefile `dropin`): they launch mgs_kernel2<<<1, (32,32)>>> / mgs_kernel<<<1, 256>>>
    themselves and call mgs_caqr_panel_256x32; the CSV files they write must hold a QR factorisation."""
    mgs, caqr = ROOT / "oracle/_ref/test_mgs_panel_b200", ROOT / "oracle/_ref/test_caqr_panel_b200"
    if not mgs.exists() or not caqr.exists():
        pytest.skip("oracle/_ref drop-in panel drivers not built (need /root/reference at build time)")
    r = subprocess.run([str(mgs)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    A, Q, R = (_read_csv(tmp_path / f) for f in ("A.csv", "Q.csv", "R.csv"))
    assert A.shape == (256, 32) and Q.shape == (256, 32) and R.shape == (32, 32)
    assert np.abs(np.tril(R, -1)).max() == 0.0 and (np.diag(R) > 0).all()
    assert np.linalg.norm(A - Q @ R) <= 3e-5 * np.linalg.norm(A)          # (6 decimals in the CSV)
    assert np.linalg.norm(Q.T @ Q - np.eye(32)) <= 1e-4
    r = subprocess.run([str(caqr)], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "mgs_caqr_panel_256x32 block takes" in r.stdout


@pytest.mark.parametrize("m,dist", [(1000, "normal"), (16384, "uniform"), (131072, "normal"),
                                    (262144, "uniform")])
def test_tensor_core_apply_vs_forward_substitution(qr, m, dist, monkeypatch):
    """Tall panels inside the recursion form Q = A R^-1 with the split-precision tcgen05 apply
    (panel_tc.cu).  Same R, and a Q that agrees with the fp32 forward-substitution apply far below
    the fp16 rounding that the recursion applies to Q next (4.9e-4)."""
    g = torch.Generator(device="cuda").manual_seed(21)
    A0 = (torch.rand if dist == "uniform" else torch.randn)(m, 128, device="cuda", generator=g)
    out = {}
    for tc in ("0", "1"):
        monkeypatch.setenv("LB_APPLY_TC", tc)
        ctx = qr.Context()                       # the knobs are read when the context is created
        A = qr.to_colmajor(A0)
        R = qr.colmajor_empty(128, 128)
        R.fill_(float("nan"))
        qr.mgs_caqr_panel_256x128(ctx, m, 128, A, m, R, 128)
        torch.cuda.synchronize()
        assert ctx.last_launch_count == (5 if tc == "1" else 4)
        out[tc] = (A, R)
        ctx.close()
    (Q0, R0), (Q1, R1) = out["0"], out["1"]
    assert torch.equal(R0, R1)
    assert (Q1 - Q0).abs().max().item() <= 4e-6 * Q0.abs().max().item()
    eye = torch.eye(128, device="cuda", dtype=torch.float64)
    orth = torch.linalg.matrix_norm(Q1.double().t() @ Q1.double() - eye).item()
    back = (torch.linalg.matrix_norm(Q1.double() @ torch.triu(R1).double() - A0.double())
            / torch.linalg.matrix_norm(A0.double())).item()
    assert orth <= 2e-5 and back <= 1e-6


@pytest.mark.parametrize("m,n,dist", [(65536, 256, "normal"), (131072, 512, "uniform"), (70000, 128, "graded")])
def test_integer_gram_vs_fp64_gram(qr, m, n, dist, monkeypatch):
    """Tall panels inside the recursion form G = A^T A on the integer tensor path (Ozaki splitting,
    tcgen05 kind::i8, exact int32 accumulation; panel_tc.cu).  Against the fp64 DMMA kernel the
    factorisation must agree to fp32 rounding, far inside the fp16-level parity tolerances."""
    g = torch.Generator(device="cuda").manual_seed(22)
    A0 = (torch.rand if dist == "uniform" else torch.randn)(m, n, device="cuda", generator=g)
    if dist == "graded":      # columns spanning four decades, some entries 1e-4 of their column's largest
        A0 = A0 * torch.logspace(0, -4, n, device="cuda")[None, :] * (1e-4 + torch.rand(m, 1, device="cuda", generator=g))
    out = {}
    for i8 in ("0", "1"):
        monkeypatch.setenv("LB_GRAM_I8", i8)
        c = qr.Context()
        A = qr.to_colmajor(A0)
        R = qr.colmajor_empty(n, n)
        qr.later_rgsqrf(c, m, n, A, m, R, n)
        torch.cuda.synchronize()
        out[i8] = (A, R, c.last_launch_count)
        c.close()
    (Q0, R0, l0), (Q1, R1, l1) = out["0"], out["1"]
    # per panel: three conditional fp64 fallback launches and the conditional forward-substitution apply on
    # top of the fp64 path's five (the integer Gram kernel replaces the fp64 one), plus one column-maxima
    # pass for the first panel (the update that produces a later panel leaves its maxima behind)
    assert l1 == l0 + 4 * (n // 128) + 1
    assert (R1 - R0).abs().max().item() <= 5e-6 * R0.abs().max().item()
    # (a last-bit change of R12 can flip fp16 roundings of the update's operands: Q moves at that level)
    assert (Q1 - Q0).abs().max().item() <= 4.9e-4 * Q0.abs().max().item()
    assert qr.orthogonality(Q1) <= 1.1 * qr.orthogonality(Q0) + 1e-8


@pytest.mark.parametrize("m,n", [(65536, 256), (131072, 1024), (70008, 512)])
def test_cast_fused_into_gram_load_is_bit_identical(qr, m, n, monkeypatch):
    """Left-spine Gram products of tall matrices read the caller's fp32 columns and round them to fp16
    inside the kernel's load path (tc_gram_cast.cu) instead of running the cast kernel first: same
    rounding, same split-K factor, same accumulation order - same bits, one launch less per node."""
    g = torch.Generator(device="cuda").manual_seed(23)
    A0 = torch.randn(m, n, device="cuda", generator=g)
    out = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("LB_GRAM_CAST", fused)
        c = qr.Context()
        A = qr.to_colmajor(A0)
        R = qr.colmajor_empty(n, n)
        R.fill_(float("nan"))
        qr.later_rgsqrf(c, m, n, A, m, R, n)
        torch.cuda.synchronize()
        out[fused] = (A, R, c.last_launch_count)
        c.close()
    (Q0, R0, l0), (Q1, R1, l1) = out["0"], out["1"]
    assert torch.equal(R0, R1) and torch.equal(Q0, Q1)
    assert l1 < l0


@pytest.mark.parametrize("m,n", [(2048, 1024), (1000, 512), (18944, 512)])
def test_fused_node_kernel_matches_the_three_launch_path(qr, m, n, monkeypatch):
    """Small nodes (half-width 128 by default, 256 with LB_NODE_FUSE=256) of matrices of at most 128 x #SMs
    rows run as ONE kernel - Gram product, fixed-order reduce over the row tiles, update - instead of three
    (tc_node_kernel, tc_update.cu).  The sums are taken over different partitions of the rows, so R agrees to
    rounding, not bit for bit; the factorisation is as accurate either way, deterministic, and shorter."""
    g = torch.Generator(device="cuda").manual_seed(29)
    A0 = torch.randn(m, n, device="cuda", generator=g)
    out = {}
    for level in ("0", "128", "256"):
        monkeypatch.setenv("LB_NODE_FUSE", level)
        c = qr.Context()
        runs = []
        for _ in range(3):                      # direct launches, graph capture, graph replay
            A = qr.to_colmajor(A0)
            R = qr.colmajor_empty(n, n)
            R.fill_(float("nan"))
            qr.later_rgsqrf(c, m, n, A, m, R, n)
            torch.cuda.synchronize()
            runs.append((A, R))
        assert all(torch.equal(runs[0][0], a) and torch.equal(runs[0][1], r) for a, r in runs[1:])
        out[level] = (runs[0][0], runs[0][1], c.last_launch_count, qr.backward_error(A0, *runs[0]),
                      qr.orthogonality(runs[0][0]))
        c.close()
    assert out["256"][2] < out["128"][2] < out["0"][2]
    R0 = out["0"][1]
    for level in ("128", "256"):
        Q, R, _, back, orth = out[level]
        assert torch.isfinite(R).all() and (torch.tril(R, -1) == 0).all()
        assert (R - R0).norm() <= 1e-3 * R0.norm()      # (fp16 roundings downstream flip: backward-error level)
        assert back <= 1.05 * out["0"][3] + 1e-7 and orth <= 1.05 * out["0"][4] + 1e-7


# ------------------------------------------------------------------------------ full-size properties
def _factor_device(qr, ctx, A0: torch.Tensor):
    m, n = A0.shape
    A = qr.to_colmajor(A0)
    R = qr.colmajor_empty(n, n)
    R.fill_(float("nan"))
    qr.later_rgsqrf(ctx, m, n, A, m, R, n)
    return A, R


@pytest.mark.parametrize("m,n,dist,back_max,orth_max", [
    (16384, 16384, "uniform", 5.1e-4, 7.9e-4),   # config 2; reference on its own input: 2.52e-4 / 3.90e-4
    (262144, 256, "normal", 1e-4, 1e-5),         # config 3
    (131072, 1024, "normal", 2e-4, 2e-5),        # one 8-GPU shard of config 4
    (32768, 32768, "normal", 4.2e-4, 2.8e-4),    # config 5; reference on the same input: 2.07e-4 / 1.35e-4
])
def test_full_size_properties(qr, ctx, m, n, dist, back_max, orth_max):
    g = torch.Generator(device="cuda").manual_seed(3000)
    A0 = (torch.rand if dist == "uniform" else torch.randn)(m, n, device="cuda", generator=g)
    Q, R = _factor_device(qr, ctx, A0)
    assert torch.isfinite(R).all()
    assert torch.tril(R, -1).abs().max().item() == 0.0          # R upper triangular, explicitly
    assert (R.diagonal() > 0).all()
    back, orth = qr.backward_error(A0, Q, R), qr.orthogonality(Q)
    from tests.helpers import reflib
    if reflib.available():
        # the bar itself: the UNMODIFIED reference on the same device buffer, at full size
        Qr, Rr = reflib.RefLib().rgsqrf(A0)
        back_ref, orth_ref = qr.backward_error(A0, Qr, Rr), qr.orthogonality(Qr)
        del Qr, Rr
        assert back <= 2 * back_ref and orth <= 2 * orth_ref, (back, back_ref, orth, orth_ref)
    assert back <= back_max                                     # A = Q R   (bounds: 2x the reference's values,
    assert orth <= orth_max                                     # Q^T Q = I  measured once, for boxes without it)
    # column norms: |r_jj| <= ||a_j||, and ||R||_F = ||A||_F up to the orthogonality defect
    assert abs(float(torch.linalg.norm(R.double()) / torch.linalg.norm(A0.double())) - 1.0) <= 1e-2


def test_scaling_by_powers_of_two(qr, ctx):
    """Linearity: QR(4 A) = Q (4 R).  Every operation commutes with power-of-two scaling except
    the fp16 casts of entries below the fp16 normal range (|x| < 6.1e-5), so the two runs agree to
    far better than the algorithm's own fp16-level accuracy, but not bit for bit."""
    g = torch.Generator(device="cuda").manual_seed(7)
    A0 = torch.randn(2048, 512, device="cuda", generator=g)
    Q1, R1 = _factor_device(qr, ctx, A0)
    Q2, R2 = _factor_device(qr, ctx, A0 * 4.0)
    assert (Q1 - Q2).abs().max().item() <= 1e-5
    assert (R1 * 4.0 - R2).abs().max().item() <= 1e-5 * R2.abs().max().item()


def test_requantising_q_is_idempotent(qr, ctx):
    """QR of an (almost) orthonormal Q returns R ~ I and Q back."""
    g = torch.Generator(device="cuda").manual_seed(8)
    Q, _ = _factor_device(qr, ctx, torch.randn(4096, 512, device="cuda", generator=g))
    Q2, R2 = _factor_device(qr, ctx, Q.clone())
    eye = torch.eye(512, device="cuda")
    assert (R2 - eye).abs().max().item() <= 2e-3
    assert (Q2 - Q).abs().max().item() <= 2e-3


def test_deterministic_and_graph_equals_stream(qr):
    g = torch.Generator(device="cuda").manual_seed(9)
    A0 = torch.rand(4096, 1024, device="cuda", generator=g)
    outs = []
    for use_graph in (True, True, False):
        c = qr.Context(use_graph=use_graph)
        for _ in range(2):                       # second call replays the cached graph
            Q, R = _factor_device(qr, c, A0)
        A2 = qr.to_colmajor(A0)
        R2 = qr.colmajor_empty(1024, 1024)
        qr.later_rgsqrf(c, 4096, 1024, A2, 4096, R2, 1024)
        outs.append((A2.clone(), R2.clone()))
        c.close()
    for Q, R in outs[1:]:
        assert torch.equal(Q, outs[0][0]) and torch.equal(R, outs[0][1])


# ------------------------------------------------------------------------------ boundary behaviour
def _host_block_mask(n):
    """Entries of hR the host entry point writes: the block upper triangle at its transfer
    granularity max(min(n, 256), n/16) (include/later_b200.h)."""
    c = max(min(n, 256), n // 16)
    blk = np.arange(n) // c
    return blk[:, None] <= blk[None, :]


@pytest.mark.parametrize("m,n,pinned,pad", [(1024, 256, True, 0), (1024, 128, True, 0),
                                            (4096, 2048, True, 0), (8192, 4096, True, 0),
                                            (2048, 512, False, 0), (2048, 512, True, 24),
                                            (65544, 256, True, 8)])    # tall: tensor-core panel kernels
def test_host_entry_point_matches_device_entry_point(qr, ctx, m, n, pinned, pad):
    """Same bits as the device entry point, on the first (direct), second (graph capture) and third
    (graph replay) call; entries below the block diagonal of hR are left alone."""
    rng = np.random.default_rng(10)
    A0 = rng.standard_normal((m, n), dtype=np.float32)
    Qd, Rd = run_rgsqrf(qr, ctx, A0)
    lda, ldr = m + pad, n + pad
    bufA = torch.empty((n, lda), dtype=torch.float32)
    bufR = torch.empty((n, ldr), dtype=torch.float32)
    if pinned:
        bufA, bufR = bufA.pin_memory(), bufR.pin_memory()
    hA, hR = bufA.t()[:m], bufR.t()[:n]
    mask = _host_block_mask(n)
    for _ in range(3):
        bufA.fill_(-3.0)
        hA.copy_(torch.from_numpy(A0))
        bufR.fill_(7.0)
        qr.later_rgsqrf_host(ctx, m, n, hA, lda, hR, ldr)
        assert np.array_equal(hA.numpy(), Qd)
        got = hR.numpy()
        assert np.array_equal(got[mask], Rd[mask])
        assert np.all(got[~mask] == 7.0)
        if pad:
            assert np.all(bufA.t()[m:].numpy() == -3.0) and np.all(bufR.t()[n:].numpy() == 7.0)


def test_stream_in_matches_device_entry_point(qr, ctx):
    """Host in, device out (the local step of the multi-GPU path): same bits as factoring a matrix
    that is already on the device, on the direct, the capturing and the replaying call."""
    m, n = 8192, 1024
    rng = np.random.default_rng(15)
    A0 = rng.standard_normal((m, n), dtype=np.float32)
    Qd, Rd = run_rgsqrf(qr, ctx, A0)
    hA = torch.empty((n, m), dtype=torch.float32).pin_memory().t()
    hA.copy_(torch.from_numpy(A0))
    A = qr.colmajor_empty(m, n)
    R = qr.colmajor_empty(n, n)
    for _ in range(3):
        A.fill_(float("nan"))
        R.fill_(float("nan"))
        qr.later_rgsqrf_stream_in(ctx, m, n, hA, m, A, m, R, n)
        torch.cuda.synchronize()
        assert np.array_equal(hA.numpy(), A0)                       # the host copy is input only
        assert np.array_equal(A.cpu().numpy(), Qd)
        assert np.array_equal(R.cpu().numpy(), Rd)                  # including the zeros below


def test_leading_dimensions_are_honoured(qr, ctx):
    rng = np.random.default_rng(11)
    A0 = rng.standard_normal((512, 256), dtype=np.float32)
    Q1, R1 = run_rgsqrf(qr, ctx, A0)
    A = qr.colmajor_empty(512, 256, ld=520)
    A.copy_(torch.from_numpy(A0))
    R = qr.colmajor_empty(256, 256, ld=300)
    qr.later_rgsqrf(ctx, 512, 256, A, 520, R, 300)
    assert np.array_equal(A.cpu().numpy(), Q1) and np.array_equal(R.cpu().numpy(), R1)


def test_leading_dimensions_are_honoured_on_the_tall_path(qr, ctx):
    """Same with the tensor-core panel kernels and the cast-fused Gram product (m >= 65536), a row
    count that is not a multiple of 128 and leading dimensions that are not the row count."""
    m, n = 65544, 512
    g = torch.Generator(device="cuda").manual_seed(24)
    A0 = torch.randn(m, n, device="cuda", generator=g)
    A1 = qr.to_colmajor(A0)
    R1 = qr.colmajor_empty(n, n)
    qr.later_rgsqrf(ctx, m, n, A1, m, R1, n)
    A2 = qr.colmajor_empty(m, n, ld=m + 40)
    A2.copy_(A0)
    R2 = qr.colmajor_empty(n, n, ld=n + 12)
    qr.later_rgsqrf(ctx, m, n, A2, m + 40, R2, n + 12)
    assert torch.equal(A1, A2) and torch.equal(R1, R2)
    assert qr.backward_error(A0, A2, R2) <= 1e-4 and qr.orthogonality(A2) <= 1e-5


def test_argument_errors_do_not_launch(qr, ctx):
    A = qr.colmajor_empty(512, 192)
    R = qr.colmajor_empty(192, 192)
    with pytest.raises(qr.LaterError) as e:
        qr.later_rgsqrf(ctx, 512, 192, A, 512, R, 192)          # n not a multiple of 128
    assert e.value.code == -1
    A = qr.colmajor_empty(512, 384)
    R = qr.colmajor_empty(384, 384)
    with pytest.raises(qr.LaterError) as e:
        qr.later_rgsqrf(ctx, 512, 384, A, 512, R, 384)          # n = 3 * 128: halving never reaches 128
    assert e.value.code == -1
    A = qr.colmajor_empty(128, 256)
    R = qr.colmajor_empty(256, 256)
    with pytest.raises(qr.LaterError):
        qr.later_rgsqrf(ctx, 128, 256, A, 128, R, 256)          # m < n


# ------------------------------------------------------------------------------ later_ormqr
def test_ormqr_vs_oracle_and_reference(qr, ctx):
    kind, m, n, dist, seed = CASES["ormqr_512x256"]
    W0, Y0 = make_input(kind, m, n, dist, seed)
    g = np.load(GOLD / "ormqr_512x256.npz")
    for fn, key, of in ((qr.later_ormqr, "Q_ormqr", orc.later_ormqr), (qr.later_ormqr2, "Q_ormqr2", orc.later_ormqr2)):
        W, Y = dev_colmajor(qr, W0), dev_colmajor(qr, Y0)
        fn(m, n, W, m, Y, m, ctxt=ctx)
        out = W.cpu().numpy()
        ref64 = _ormqr_fp64(W0, Y0, merge=(key == "Q_ormqr"))
        scale = np.abs(ref64).max()
        # fp32-faithful: as close to the fp64 result as the reference's own fp32 cuBLAS path is
        err_ref = np.abs(g[key].astype(np.float64) - ref64[::8, :]).max()
        assert np.abs(out - ref64).max() <= max(4 * err_ref, 4e-6 * scale)
        assert np.abs(out[::8, :] - g[key]).max() <= 1e-5 * scale


def _ormqr_fp64(W0, Y0, merge):
    W, Y = W0.astype(np.float64).copy(), Y0.astype(np.float64)
    m, n = W.shape
    h = n // 2
    if merge:
        W[:, h:] -= W[:, :h] @ (Y[:, :h].T @ W[:, h:])
    return np.eye(m, n) - W @ Y[:n, :n].T


def test_ormqr_large_is_fp32_faithful(qr, ctx):
    m, n = 4096, 1024
    g = torch.Generator(device="cuda").manual_seed(12)
    Y0 = torch.tril(torch.randn(m, n, device="cuda", generator=g) * 0.05, -1)
    Y0.diagonal().fill_(1.0)
    W0 = torch.randn(m, n, device="cuda", generator=g) * 0.05
    W, Y = qr.to_colmajor(W0), qr.to_colmajor(Y0)
    qr.later_ormqr2(m, n, W, m, Y, m, ctxt=ctx)
    ref = torch.eye(m, n, device="cuda", dtype=torch.float64) - W0.double() @ Y0.double()[:n, :n].t()
    err = (W.double() - ref).abs().max().item() / ref.abs().max().item()
    err32 = ((torch.eye(m, n, device="cuda") - W0 @ Y0[:n, :n].t()).double() - ref).abs().max().item() / ref.abs().max().item()
    assert err <= max(4 * err32, 2e-6)


def test_ormqr_long_accumulation_is_fp32_faithful(qr, ctx):
    """K = 8192: the tensor core's truncating fp32 accumulate would cost a digit here (and more at
    config 5's K = 32768) unless the hi*hi term is accumulated in chunks; the full later_ormqr
    (merge step + I - W Y^T) must stay at the level of an fp32 FMA implementation."""
    m = n = 8192
    h = n // 2
    g = torch.Generator(device="cuda").manual_seed(14)
    Y0 = torch.tril(torch.randn(m, n, device="cuda", generator=g) * 0.02, -1)
    Y0.diagonal().fill_(1.0)
    W0 = torch.randn(m, n, device="cuda", generator=g) * 0.02
    W, Y = qr.to_colmajor(W0), qr.to_colmajor(Y0)
    qr.later_ormqr(m, n, W, m, Y, m, ctxt=ctx)

    def form(Wm, Ym):
        Wm = Wm.clone()
        Wm[:, h:] -= Wm[:, :h] @ (Ym[:, :h].t() @ Wm[:, h:])
        return torch.eye(m, n, device="cuda", dtype=Wm.dtype) - Wm @ Ym[:n, :n].t()

    ref = form(W0.double(), Y0.double())
    scale = ref.abs().max().item()
    err = (W.double() - ref).abs().max().item() / scale
    err32 = (form(W0, Y0).double() - ref).abs().max().item() / scale     # fp32 cuBLAS, TF32 off
    assert err <= max(2 * err32, 1e-6)


# ------------------------------------------------------------------------------ TSQR back-multiply
def test_tsqr_apply(qr):
    c = qr.Context()
    g = torch.Generator(device="cuda").manual_seed(13)
    A0 = torch.randn(8192, 256, device="cuda", generator=g)
    Q, R = _factor_device(qr, c, A0)
    W0 = torch.linalg.qr(torch.randn(256, 256, device="cuda", generator=g))[0]
    W = qr.to_colmajor(W0)
    Qbefore = Q.clone()
    qr.tsqr_apply(c, 8192, 256, Q, 8192, W, 256)
    ref = Qbefore.half().float() @ W0.half().float()
    assert (Q - ref).abs().max().item() <= 2e-5 * max(1.0, ref.abs().max().item())
    c.close()


# ------------------------------------------------------------------------------ the reference's own driver
def _parse_driver(out: str):
    import re
    orth = float(re.search(r"\|\|I-Q'\*Q\|\|/N = ([0-9.eE+-]+)", out).group(1))
    back = float(re.search(r"Backward error: \|\|A-QR\|\|/\(\|\|A\|\|\) = ([0-9.eE+-]+)", out).group(1))
    return back, orth


@pytest.mark.parametrize("algo,m,n", [(1, 1024, 1024), (1, 4096, 2048), (2, 2048, 1024), (3, 2048, 1024)])
def test_reference_driver_runs_unchanged_against_this_library(algo, m, n):
    """test/test_qr.cu of the reference, compiled unmodified: once with the reference's own sources
    (test_qr_ref) and once against later_b200 (test_qr_b200); same cuRAND input (seed 3000)."""
    ref_bin, new_bin = ROOT / "oracle/_ref/test_qr_ref", ROOT / "oracle/_ref/test_qr_b200"
    if not ref_bin.exists() or not new_bin.exists():
        pytest.skip("oracle/_ref binaries not built (need /root/reference at build time)")
    outs = []
    for exe in (ref_bin, new_bin):
        r = subprocess.run([str(exe), str(algo), str(m), str(n), "-check"], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        outs.append(_parse_driver(r.stdout))
    (back_ref, orth_ref), (back_new, orth_new) = outs
    if algo == 1:
        assert back_new <= 2 * back_ref and orth_new <= 2 * orth_ref, outs
    else:
        # algo 2 / 3: later_rhouqr + later_ormqr / later_bhouqr + later_ormqr2.  (The reference's current
        # later_rhouqr merges the top-level W block itself AND its driver's later_ormqr merges it again -
        # the guard at QR/later_rhouqr.cu:165 is commented out - so its own numbers may be off; the bar
        # here is absolute as well.)
        assert back_new <= max(2 * back_ref, 1e-3) and orth_new <= max(2 * orth_ref, 1e-3), outs
        assert np.isfinite(back_new) and np.isfinite(orth_new)
