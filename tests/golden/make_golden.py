"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/_ref/libref_later.so,
built by oracle/Makefile from /root/reference) on a B200.

    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden'   # then copy *.npz here

Inputs are regenerated from the seeds below (numpy PCG64), so only outputs are stored: the full R
factor and every 8th row of Q (enough to pin a CPU restatement; keeps the fixtures small).
"""
import ctypes as C
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]

CASES = {
    # name: (kind, m, n, distribution, seed)
    "panel_1024x128_uniform": ("panel", 1024, 128, "uniform", 11),
    "panel_832x128_normal": ("panel", 832, 128, "normal", 12),       # m % 256 = 64: remainder path
    "rgsqrf_512x256_normal": ("rgsqrf", 512, 256, "normal", 13),
    "rgsqrf_768x256_uniform": ("rgsqrf", 768, 256, "uniform", 14),
    "rgsqrf_1024x512_normal": ("rgsqrf", 1024, 512, "normal", 15),
    "ormqr_512x256": ("ormqr", 512, 256, "wy", 16),
    # 32-column entry points (reference QR/panel.cu:65-134, :246-325)
    "panel32_4096x32_uniform": ("panel32", 4096, 32, "uniform", 17),
    "panel32_1000x32_normal": ("panel32", 1000, 32, "normal", 18),   # m % 256 = 232: remainder block
    "mgs2_700x32_normal": ("mgs2", 700, 32, "normal", 19),           # three blocks, the last one ragged
    # conditioning sweep on the 128-column panel: A = G diag(sigma) V^T, sigma log-spaced 1 .. 1/kappa
    **{f"panel_4096x128_cond1e{k}": ("panel", 4096, 128, f"cond1e{k}", 30 + k) for k in range(1, 7)},
}

# Conditioning sweep at sizes where only the metrics are kept (golden_meta.json["sweep"]): inputs
# come from tests/helpers/inputs.py (torch CUDA generator), so they exist on the GPU box only.
SWEEP_SHAPES = [(131072, 128), (16384, 1024), (131072, 1024)]
SWEEP_KAPPAS = [1e1, 1e2, 1e3, 1e4, 1e5, 1e6]


def make_input(kind, m, n, dist, seed):
    rng = np.random.default_rng(seed)
    if dist == "uniform":
        return np.asfortranarray(rng.random((m, n), dtype=np.float32))
    if dist == "normal":
        return np.asfortranarray(rng.standard_normal((m, n), dtype=np.float32))
    if dist.startswith("cond"):   # kappa(A) ~ the requested value (G is only nearly orthonormal)
        kappa = float(dist[4:])
        G = rng.standard_normal((m, n)) / np.sqrt(m)
        V, _ = np.linalg.qr(rng.standard_normal((n, n)))
        sig = np.logspace(0.0, -np.log10(kappa), n)
        return np.asfortranarray(((G * sig) @ V.T).astype(np.float32))
    if dist == "wy":  # a WY-like pair: Y unit lower trapezoidal, W of similar magnitude
        Y = np.tril(rng.standard_normal((m, n), dtype=np.float32) * 0.1, -1)
        Y[np.arange(n), np.arange(n)] = 1.0
        W = (rng.standard_normal((m, n), dtype=np.float32) * 0.1).astype(np.float32)
        return np.asfortranarray(W), np.asfortranarray(Y)
    raise ValueError(dist)


def main(out_dir):
    import torch
    out = Path(out_dir)
    out.mkdir(parents=True, exist_ok=True)
    ref = C.CDLL(str(ROOT / "oracle" / "_ref" / "libref_later.so"))
    vp, ci = C.c_void_p, C.c_int
    ref.ref_later_rgsqrf.argtypes = [ci, ci, vp, ci, vp, ci, vp, ci, vp, ci]
    ref.ref_mgs_caqr_panel_256x128.argtypes = [ci, ci, vp, ci, vp, ci, vp]
    ref.ref_later_ormqr.argtypes = [ci, ci, vp, ci, vp, ci, vp]
    ref.ref_later_ormqr2.argtypes = [ci, ci, vp, ci, vp, ci, vp]
    ref.ref_mgs_caqr_panel_256x32.argtypes = [ci, ci, vp, ci, vp, ci, vp]
    ref.ref_mgs_kernel2.argtypes = [ci, ci, vp, ci, vp, ci]

    def dev(a):  # column-major numpy -> device tensor with the same memory layout
        return torch.from_numpy(np.ascontiguousarray(a.T)).cuda()

    def host(t, m, n):
        return np.asfortranarray(t.cpu().numpy().T.reshape(m, n, order="A"))

    meta = {}
    for name, (kind, m, n, dist, seed) in CASES.items():
        if kind == "mgs2":       # every 256-row block on its own; R_b stacked at rows 32 b
            A0 = make_input(kind, m, n, dist, seed)
            nb = (m + 255) // 256
            dA = dev(A0)
            dR = torch.zeros((n, nb * 32), device="cuda", dtype=torch.float32)     # (nb*32) x n, ld = nb*32
            rc = ref.ref_mgs_kernel2(m, n, dA.data_ptr(), m, dR.data_ptr(), nb * 32)
            torch.cuda.synchronize()
            assert rc == 0, (name, rc)
            np.savez_compressed(out / f"{name}.npz", R=dR.cpu().numpy().T, Q=dA.cpu().numpy().T)
            meta[name] = dict(kind=kind, m=m, n=n, dist=dist, seed=seed)
        elif kind in ("panel", "rgsqrf", "panel32"):
            A0 = make_input(kind, m, n, dist, seed)
            dA = dev(A0)
            dR = torch.zeros((n, n), device="cuda", dtype=torch.float32)
            work = torch.zeros(8 * m * max(n, 32) // 8 + 65536, device="cuda", dtype=torch.float32)
            hwork = torch.zeros(m * n, device="cuda", dtype=torch.float16)
            if kind == "panel32":
                rc = ref.ref_mgs_caqr_panel_256x32(m, n, dA.data_ptr(), m, dR.data_ptr(), n, work.data_ptr())
            elif kind == "panel":
                rc = ref.ref_mgs_caqr_panel_256x128(m, n, dA.data_ptr(), m, dR.data_ptr(), n, work.data_ptr())
            else:
                rc = ref.ref_later_rgsqrf(m, n, dA.data_ptr(), m, dR.data_ptr(), n, work.data_ptr(),
                                          work.numel(), hwork.data_ptr(), hwork.numel())
            torch.cuda.synchronize()
            assert rc == 0, (name, rc)
            Q = dA.cpu().numpy().T            # (m, n)
            R = dR.cpu().numpy().T
            A64, Q64, R64 = A0.astype(np.float64), Q.astype(np.float64), R.astype(np.float64)
            back = np.linalg.norm(A64 - Q64 @ R64) / np.linalg.norm(A64)
            orth = np.linalg.norm(np.eye(n) - Q64.T @ Q64) / n
            np.savez_compressed(out / f"{name}.npz", R=R, Q_rows8=Q[::8, :].copy(),
                                backward=np.float64(back), orth=np.float64(orth))
            meta[name] = dict(kind=kind, m=m, n=n, dist=dist, seed=seed, backward=back, orth=orth)
            if dist.startswith("cond"):
                meta[name]["cond"] = float(np.linalg.cond(A64))
        else:
            W0, Y0 = make_input(kind, m, n, dist, seed)
            res = {}
            for fn, key in ((ref.ref_later_ormqr, "Q_ormqr"), (ref.ref_later_ormqr2, "Q_ormqr2")):
                dW, dY = dev(W0), dev(Y0)
                work = torch.zeros(m * n, device="cuda", dtype=torch.float32)
                rc = fn(m, n, dW.data_ptr(), m, dY.data_ptr(), m, work.data_ptr())
                torch.cuda.synchronize()
                assert rc == 0, (name, rc)
                res[key] = dW.cpu().numpy().T[::8, :].copy()
            np.savez_compressed(out / f"{name}.npz", **res)
            meta[name] = dict(kind=kind, m=m, n=n, dist=dist, seed=seed)
        print(name, meta[name], flush=True)
    # conditioning sweep at full size: the reference's two metrics per (shape, kappa)
    sys.path.insert(0, str(ROOT))
    from tests.helpers.inputs import cond_matrix, metrics
    from tests.helpers.reflib import RefLib
    rl = RefLib()
    sweep = {}
    for (m, n) in SWEEP_SHAPES:
        for kappa in SWEEP_KAPPAS:
            A0 = cond_matrix(m, n, kappa, seed=4000 + n)
            Q, R = rl.rgsqrf(A0)
            back, orth = metrics(A0, Q, R)
            sweep[f"{m}x{n}_cond{kappa:.0e}"] = dict(m=m, n=n, kappa=kappa, backward=back, orth=orth)
            print("sweep", m, n, kappa, back, orth, flush=True)
            del A0, Q, R
    meta["sweep"] = sweep
    (out / "golden_meta.json").write_text(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parent))
