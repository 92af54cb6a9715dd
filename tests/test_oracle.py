"""The oracle (oracle/rgsqrf_oracle.py, a numpy restatement of the reference) against
 (1) outputs of the reference itself, generated on a B200 by tests/golden/make_golden.py from
     oracle/_ref/libref_later.so and committed as fixtures - this is what pins the oracle;
 (2) host LAPACK sgeqrf, the CPU stand-in BASELINE.json names;
 (3) the structural facts the reference guarantees (R upper triangular, r_ii > 0, edge shapes)."""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from oracle import rgsqrf_oracle as orc  # noqa: E402
from tests.golden.make_golden import CASES, make_input  # noqa: E402

GOLD = ROOT / "tests" / "golden"
META = json.loads((GOLD / "golden_meta.json").read_text()) if (GOLD / "golden_meta.json").exists() else {}


def _need(name):
    if name not in META or not (GOLD / f"{name}.npz").exists():
        pytest.fail(f"golden fixture {name} missing: run tests/golden/make_golden.py on the GPU box")
    return np.load(GOLD / f"{name}.npz")


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v[0] in ("panel", "rgsqrf")])
def test_oracle_matches_reference_outputs(name):
    kind, m, n, dist, seed = CASES[name]
    g = _need(name)
    A0 = make_input(kind, m, n, dist, seed)
    if kind == "panel":
        Q = np.array(A0, dtype=np.float32, order="F", copy=True)
        R = np.zeros((n, n), dtype=np.float32)
        orc.mgs_caqr_panel_256x128(Q, R)
        tol = 2e-5          # pure fp32 path: differences are summation order only
    else:
        Q, R = orc.later_rgsqrf(A0)
        tol = 2e-3          # fp16-rounded operands: a 1-ulp fp16 flip moves entries by ~5e-4 relative
    # two fp32 MGS runs with different summation orders differ by ~ eps * cond: the sweep cases
    # carry their condition number
    tol *= max(1.0, META[name].get("cond", 1.0) / 10.0)
    scale = np.abs(g["R"]).max()
    if tol < 0.1:
        assert np.abs(np.triu(R) - np.triu(g["R"])).max() <= tol * scale
        assert np.abs(Q[::8, :] - g["Q_rows8"]).max() <= tol * max(1.0, np.abs(g["Q_rows8"]).max())
    # the driver's own metrics agree with the reference's to well within the 2x parity bar
    back, orth = orc.check_result(A0, Q, R), orc.check_otho(Q)
    assert back <= 1.5 * float(g["backward"]) + 1e-7
    assert orth <= 1.5 * float(g["orth"]) + 1e-7


@pytest.mark.parametrize("name", [k for k, v in CASES.items() if v[0] in ("panel32", "mgs2")])
def test_oracle_32_column_entry_points_match_reference_outputs(name):
    """mgs_caqr_panel_256x32 (CAQR tree, QR/panel.cu:65-134) and mgs_kernel2 (one MGS per 256-row
    block, QR/panel.cu:246-325) against the reference's own outputs."""
    kind, m, n, dist, seed = CASES[name]
    g = _need(name)
    A0 = make_input(kind, m, n, dist, seed)
    Q = np.array(A0, dtype=np.float32, order="F", copy=True)
    if kind == "panel32":
        R = orc.mgs_caqr_panel_256x32(Q)
        assert np.abs(R - g["R"]).max() <= 2e-5 * np.abs(g["R"]).max()
        assert np.abs(Q[::8, :] - g["Q_rows8"]).max() <= 2e-5
    else:
        nb = (m + 255) // 256
        for b in range(nb):
            blk = Q[b * 256:min(m, (b + 1) * 256), :]
            Rb = orc.mgs_kernel2(blk)
            assert np.abs(Rb - g["R"][b * 32:(b + 1) * 32, :]).max() <= 2e-5 * np.abs(g["R"]).max()
        assert np.abs(Q - g["Q"]).max() <= 2e-5


def test_oracle_ormqr_matches_reference_outputs():
    kind, m, n, dist, seed = CASES["ormqr_512x256"]
    g = _need("ormqr_512x256")
    W0, Y0 = make_input(kind, m, n, dist, seed)
    q1 = orc.later_ormqr(W0, Y0)
    q2 = orc.later_ormqr2(W0, Y0)
    assert np.abs(q1[::8, :] - g["Q_ormqr"]).max() <= 1e-5 * max(1.0, np.abs(g["Q_ormqr"]).max())
    assert np.abs(q2[::8, :] - g["Q_ormqr2"]).max() <= 1e-5 * max(1.0, np.abs(g["Q_ormqr2"]).max())


@pytest.mark.parametrize("m,n,dist", [(512, 256, "normal"), (1024, 512, "uniform"), (384, 128, "normal")])
def test_oracle_vs_lapack(m, n, dist):
    rng = np.random.default_rng(5)
    A = rng.standard_normal((m, n), dtype=np.float32) if dist == "normal" else rng.random((m, n), dtype=np.float32)
    Q, R = orc.later_rgsqrf(A)
    Ql, Rl = orc.lapack_qr(A)
    # fp16-tensor-core tolerance vs LAPACK: backward error <= ~2 u_fp16
    assert orc.check_result(A, Q, R) <= 1e-3
    assert orc.check_result(A, Ql, Rl) <= 5e-6
    # same factorisation up to rounding: diag(R) > 0 in both, R agrees at fp16 level scaled by cond
    assert (np.diag(R) > 0).all()
    cond = np.linalg.cond(A.astype(np.float64))
    assert np.abs(R - Rl).max() <= 2e-3 * cond * np.abs(Rl).max()


def test_oracle_structure_and_edge_shapes():
    rng = np.random.default_rng(6)
    # single-block panel (m == 256), ragged panel (m % 256 != 0), smallest legal square
    for m, n in [(256, 128), (600, 128), (128, 128), (256, 256)]:
        A = rng.standard_normal((m, n), dtype=np.float32)
        Q, R = orc.later_rgsqrf(A)
        assert Q.shape == (m, n) and R.shape == (n, n)
        assert np.abs(np.tril(R, -1)).max() == 0.0
        assert (np.diag(R) > 0).all()
        assert orc.check_result(A, Q, R) < 1e-3
    with pytest.raises(ValueError):
        orc.later_rgsqrf(rng.standard_normal((256, 384), dtype=np.float32))   # n not 128*2^k, m < n
    with pytest.raises(ValueError):
        orc.later_rgsqrf(rng.standard_normal((100, 128), dtype=np.float32))   # m < n


def test_fp16_cast_is_round_to_nearest_even():
    x = np.array([1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 65504.0, 1e-8], dtype=np.float32)
    h = orc.s2h(x).astype(np.float32)
    assert h[0] == 1.0 and h[1] == np.float32(1.0 + 2.0 ** -9) and h[2] == 65504.0 and h[3] == 0.0


def test_oracle_qdwh_polar_converges_to_the_polar_factor():
    """later_qdwh_polar (EVD/later_qdwh_polar.cu:24-110) restated: on the driver's kind of input
    (symmetric U(0,1], test/test_qdwh_polar.cu:53-59) the iteration must reach the orthogonal polar
    factor to the fp16 level of the factorisations it is built on, in a handful of steps."""
    from scipy.linalg import polar
    rng = np.random.default_rng(7)
    n = 128
    H = rng.random((n, n), dtype=np.float32)
    H = 0.5 * (H + H.T)
    U, iters = orc.later_qdwh_polar(H)
    Ue, _ = polar(H.astype(np.float64))
    assert 2 <= iters <= 10
    assert np.abs(U - Ue).max() <= 5e-2
    assert np.linalg.norm(U.astype(np.float64).T @ U - np.eye(n)) / n <= 1e-4
    assert np.abs(U - U.T).max() == 0.0


@pytest.mark.parametrize("m,n", [(300, 32), (512, 128), (1024, 512)])
def test_oracle_rhouqr_is_a_householder_factorisation(m, n):
    """later_rhouqr (QR/later_rhouqr.cu:21-277) restated: Y unit lower trapezoidal, Q = I - W Y^T orthogonal
    once later_ormqr has merged the two halves of W (QR/later_ormqr.cu:27-45), A = Q R."""
    rng = np.random.default_rng(8)
    A = rng.random((m, n), dtype=np.float32)
    Y, W, R = orc.later_rhouqr(A)
    assert np.abs(np.triu(Y[:n], 1)).max() == 0.0 and np.all(np.diag(Y) == 1.0)
    assert np.abs(np.tril(R, -1)).max() == 0.0
    Q = orc.later_ormqr(W, Y) if n >= 64 else orc.later_ormqr2(W, Y)
    tol = 1e-3 if n > 256 else 2e-5          # fp16 products only where n/2 > 128 (QR/later_rhouqr.cu:83)
    assert orc.check_result(A, Q, R) <= tol
    assert orc.check_otho(Q) <= tol
