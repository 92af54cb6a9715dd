#!/usr/bin/env python
"""RGSQRF benchmark (BASELINE.json's metric): TFLOPS = (2 m n^2 - 2/3 n^3) / t, the formula the
reference prints (reference test/test_qr.cu:85-86).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1 : 16384 x 16384 on one B200 (BASELINE.json configs[1]).  The line also carries a `configs`
        block with the other single-GPU configurations (1024^2 on the reference's own cuRAND input,
        262144 x 256, 1048576 x 1024, 32768^2 + later_ormqr), each measured in the same run.
N > 1 : 1048576 x 1024, row-sharded over N GPUs (configs[3]); launched by torch.distributed.run, one
        rank per GPU, NCCL.  Every rank runs the same recursion on its row block with the panel Gram
        matrices and the R12 blocks all-reduced (later_b200_rgsqrf_dist); the classical TSQR variant
        (local QR, all-gather of R, stack QR, back-multiply) is timed beside it.  Strong scaling.  Every
        rank generates ITS ROWS OF THE SAME GLOBAL MATRIX (a generator keyed by the global element
        index), so the N-GPU factorisation is checked globally (backward error, orthogonality,
        all-reduced) and compared with a 1-GPU run of the very same matrix made by rank 0 in-run.
One JSON line on stdout (rank 0).  Inputs are synthetic N(0,1), resident in HBM before the timed
region; the matrices (1-4 GiB) are larger than the 126 MB L2, so no explicit L2 flush is needed.

--impl reference times the UNMODIFIED reference (oracle/_ref/libref_later.so: the reference's CUDA
sources compiled for sm_100 against cuBLAS 12.9 by oracle/Makefile) on the same workloads.  The
reference has no CPU path; host LAPACK sgeqrf+sorgqr is the CPU stand-in BASELINE.json names and is
reported as `cpu_baseline` on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))


def credited_flops(m: int, n: int) -> float:
    return 2.0 * n * n * (m - n / 3.0)


def algorithmic_bytes(m: int, n: int) -> float:
    """Read A once, write Q once, write R."""
    return 8.0 * m * n + 4.0 * n * n


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); smax.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------- CPU baseline
def lapack_baseline(n_sample: int = 8192) -> dict:
    """Host LAPACK sgeqrf + sorgqr (explicit Q) on a bounded sample, all host cores."""
    import numpy as np
    from scipy.linalg import lapack
    cores = os.cpu_count() or 1
    rng = np.random.default_rng(3000)
    a = np.asfortranarray(rng.standard_normal((n_sample, n_sample), dtype=np.float32))
    t0 = time.perf_counter()
    qr_, tau, _, info = lapack.sgeqrf(a, overwrite_a=1)
    q, _, info2 = lapack.sorgqr(qr_, tau, overwrite_a=1)
    dt = time.perf_counter() - t0
    assert info == 0 and info2 == 0
    return {"value": credited_flops(n_sample, n_sample) / dt / 1e12, "unit": "TFLOPS", "cores": cores,
            "kind": "port",
            "sample": f"host LAPACK sgeqrf+sorgqr (scipy OpenBLAS) on a {n_sample}x{n_sample} N(0,1) sample of the "
                      f"16384^2 workload (not extrapolated: the rate of the sample itself), {dt:.1f} s; "
                      f"stands in for a CPU path the reference does not have"}


def measured_peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# --------------------------------------------------------------------------------------- inputs
def keyed_normal(torch, m_glob: int, n: int, row0: int, rows: int, seed: int):
    """Rows [row0, row0 + rows) of the m_glob x n matrix whose element (i, j) is a standard normal
    deviate that depends only on (seed, i + j * m_glob): a counter-based generator (SplitMix64
    finaliser of the global element index, Box-Muller), so that a row shard on any rank equals those
    rows of the single-GPU matrix.  Column-major (rows x n, ld = rows)."""
    dev = "cuda"
    out = torch.empty((n, rows), device=dev, dtype=torch.float32)
    mask = (1 << 63) - 1

    def mix(z):                                   # SplitMix64 finaliser; int64 arithmetic wraps
        z = (z ^ ((z >> 30) & ((1 << 34) - 1))) * -4658895280553007687      # 0xBF58476D1CE4E5B9
        z = (z ^ ((z >> 27) & ((1 << 37) - 1))) * -7723592293110705685      # 0x94D049BB133111EB
        return z ^ ((z >> 31) & ((1 << 33) - 1))

    ii = torch.arange(rows, device=dev, dtype=torch.int64) + row0
    cols = max(1, (1 << 24) // rows)
    for j0 in range(0, n, cols):
        jj = torch.arange(j0, min(n, j0 + cols), device=dev, dtype=torch.int64)
        idx = ii[None, :] + jj[:, None] * m_glob
        a = mix(idx * 2 + seed * 0x9E3779B9 + 1)
        b = mix(idx * 2 + seed * 0x9E3779B9 + 2)
        u1 = ((a & mask) >> 10).to(torch.float64) * (1.0 / (1 << 53)) + (0.5 / (1 << 53))
        u2 = ((b & mask) >> 10).to(torch.float64) * (1.0 / (1 << 53))
        out[j0:j0 + cols] = (torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)).float()
    return out.t()


def chunked_metrics(torch, A0, Q, R):
    """(||A0 - Q R||_F^2, ||A0||_F^2, Q^T Q) of a row block, evaluated in fp64 in bounded chunks."""
    m, n = A0.shape
    Rd = torch.triu(R.double())
    res2 = torch.zeros((), device="cuda", dtype=torch.float64)
    nrm2 = torch.zeros((), device="cuda", dtype=torch.float64)
    G = torch.zeros(n, n, device="cuda", dtype=torch.float64)
    step = max(1, (1 << 26) // n)
    for r0 in range(0, m, step):
        q = Q[r0:r0 + step].double()
        a = A0[r0:r0 + step].double()
        res2 += torch.linalg.norm(a - q @ Rd) ** 2
        nrm2 += torch.linalg.norm(a) ** 2
        G += q.t() @ q
    return res2, nrm2, G


# --------------------------------------------------------------------------------------- reference arm
class ReferenceLib:
    def __init__(self):
        path = ROOT / "oracle" / "_ref" / "libref_later.so"
        if not path.exists():
            raise FileNotFoundError(f"{path} not built (oracle/Makefile needs /root/reference)")
        self.lib = C.CDLL(str(path))
        vp, ci = C.c_void_p, C.c_int
        self.lib.ref_later_rgsqrf.argtypes = [ci, ci, vp, ci, vp, ci, vp, ci, vp, ci]
        self.lib.ref_later_rgsqrf.restype = ci
        self.lib.ref_later_ormqr.argtypes = [ci, ci, vp, ci, vp, ci, vp]
        self.lib.ref_generate_uniform.argtypes = [vp, ci, ci]
        self.lib.ref_generate_uniform.restype = None

    def rgsqrf(self, m, n, A, R, work, hwork):
        rc = self.lib.ref_later_rgsqrf(m, n, A.data_ptr(), m, R.data_ptr(), n, work.data_ptr(),
                                       work.numel(), hwork.data_ptr(), hwork.numel())
        if rc != 0:
            raise RuntimeError(f"reference later_rgsqrf: cuda error {rc}")

    def ormqr(self, m, n, W, Y, work):
        rc = self.lib.ref_later_ormqr(m, n, W.data_ptr(), m, Y.data_ptr(), m, work.data_ptr())
        if rc != 0:
            raise RuntimeError(f"reference later_ormqr: cuda error {rc}")

    def uniform(self, A, m, n):
        self.lib.ref_generate_uniform(A.data_ptr(), m, n)


def bind_to_gpu_numa_node(torch, index):
    """Run this process (and allocate its pinned buffers) on the CPUs next to its GPU: with one
    rank per GPU on a two-socket box, half the ranks otherwise copy across the socket link."""
    try:
        import pynvml
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(index)
        try:
            bus = f"{pr.pci_domain_id:08x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:  # noqa: BLE001
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
    except Exception:  # noqa: BLE001
        pass


def timed(torch, fn, restore, reps=5, warm=3):
    """Median device time (ms) of fn() over `reps` runs after `warm` untimed ones (direct launch,
    graph capture, first replay); restore() puts the in-place input back, outside the event pair."""
    ts = []
    for i in range(warm + reps):
        restore()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        if i >= warm:
            ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


# --------------------------------------------------------------------------------------- configs block
def run_configs(torch, impl, qr, ctx, ref, peaks, src, skip_big):
    """The other single-GPU configurations of BASELINE.json, measured in this run: C1 (1024^2 on the
    reference driver's own input), C3 (262144 x 256), C4 on one GPU (1048576 x 1024) and C5
    (32768^2 RGSQRF + later_ormqr)."""
    out = {}

    def one(name, m, n, make_input, reps=5, accuracy=True):
        A0 = make_input(m, n)
        A = torch.empty((n, m), device="cuda").t()
        R = torch.zeros((n, n), device="cuda").t()
        if impl == "b200":
            def fn():
                qr.later_rgsqrf(ctx, m, n, A, m, R, n)
        else:
            work = torch.zeros(max(m // 256 * 32 * n, 1 << 20) + 4 * m + (1 << 20), device="cuda")
            hwork = torch.zeros(m * n, device="cuda", dtype=torch.float16)

            def fn():
                ref.rgsqrf(m, n, A, R, work, hwork)
        ms = timed(torch, fn, lambda: A.copy_(A0), reps=reps)
        entry = {"m": m, "n": n, "ms": ms, "tflops": credited_flops(m, n) / (ms * 1e-3) / 1e12,
                 "algorithmic_gbs": algorithmic_bytes(m, n) / (ms * 1e-3) / 1e9}
        entry["hbm_roofline_frac"] = entry["algorithmic_gbs"] / peaks["hbm_gbs"]
        entry["tensor_roofline_frac"] = entry["tflops"] / peaks["bf16_tflops_sustained"]
        if accuracy:
            res2, nrm2, G = chunked_metrics(torch, A0, A, R)
            G.diagonal().sub_(1.0)
            entry["backward_error"] = float(torch.sqrt(res2 / nrm2))
            entry["orthogonality"] = float(torch.linalg.norm(G)) / n
        out[name] = entry
        del A0, A, R
        torch.cuda.empty_cache()

    gen = torch.Generator(device="cuda").manual_seed(3001)

    def normal(m, n):
        return torch.empty((n, m), device="cuda").normal_(generator=gen).t()

    def curand_uniform(m, n):
        """The reference driver's input: cuRAND XORWOW, seed 3000, U(0,1] (reference util/util.cu:102-109).
        Both libraries export the same generator call; each arm uses its own."""
        buf = torch.empty((n, m), device="cuda")
        if impl == "b200":
            from later_b200._lib import lib as _l
            fn = getattr(_l, "_Z21generateUniformMatrixPfii")
            fn.argtypes = [C.c_void_p, C.c_int, C.c_int]
            fn.restype = None
            fn(buf.data_ptr(), m, n)
        else:
            ref.uniform(buf, m, n)
        torch.cuda.synchronize()
        return buf.t()

    one("C1_test_qr_1024x1024_curand_uniform", 1024, 1024, curand_uniform, reps=9)
    one("C3_panel_262144x256", 262144, 256, normal, reps=9)
    if not skip_big:
        one("C4_1048576x1024_one_gpu", 1048576, 1024, normal)
        one("C5_32768x32768", 32768, 32768, normal, reps=3)
        # later_ormqr at config 5's size on a synthetic WY pair (fp32-faithful path; reference: fp32 cuBLAS)
        m = n = 32768
        Y = torch.empty((n, m), device="cuda").normal_(generator=gen).mul_(0.01).t()
        W0 = torch.empty((n, m), device="cuda").normal_(generator=gen).mul_(0.01).t()
        W = torch.empty((n, m), device="cuda").t()
        if impl == "b200":
            def fn():
                qr.later_ormqr(m, n, W, m, Y, m, ctxt=ctx)
        else:
            work = torch.zeros((n // 2) * (n // 2), device="cuda")

            def fn():
                ref.ormqr(m, n, W, Y, work)
        ms = timed(torch, fn, lambda: W.copy_(W0), reps=2, warm=1)
        out["C5_later_ormqr_32768x32768"] = {"m": m, "n": n, "ms": ms,
                                             "tflops_3mn2": 3.0 * m * n * n / (ms * 1e-3) / 1e12}
        del Y, W0, W
        torch.cuda.empty_cache()
    out["peaks"] = {"hbm_gbs": peaks["hbm_gbs"], "bf16_tflops_sustained": peaks["bf16_tflops_sustained"],
                    "source": f"{src} MEASURED_PEAKS.json"}
    return out


# --------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--m", type=int, default=0, help="override rows (diagnostics only)")
    ap.add_argument("--n", type=int, default=0, help="override cols (diagnostics only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (diagnostics)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N = max(args.gpus, world)

    if args.impl == "reference" and rank != 0:
        return 0                                  # the reference is single-GPU: rank 0 alone runs it
    if not torch.cuda.is_available():
        if args.impl == "reference":
            print(json.dumps({"impl": "reference", "unavailable": "no CUDA device"}))
            return 0
        raise SystemExit("bench.py needs a B200: later_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    bind_to_gpu_numa_node(torch, local_rank)
    torch.backends.cuda.matmul.allow_tf32 = False
    distributed = world > 1 and args.impl == "b200"
    if distributed:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"      # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    if N == 1:
        m, n = 16384, 16384
        workload = "rgsqrf_16384x16384_square"
    else:
        m, n = 1048576, 1024
        workload = f"rgsqrf_1048576x1024_rowsharded_dp{N}"
    if args.m and args.n:
        m, n = args.m, args.n
        workload = f"rgsqrf_{m}x{n}_override"
    shards = world if distributed else 1
    m_loc = m // shards
    flops = credited_flops(m, n)
    peaks, peak_src = measured_peaks()

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    SEED = 3000
    # this rank's rows of the global matrix (N = 1, or the single-GPU reference arm: all of them)
    A0 = keyed_normal(torch, m, n, (rank if distributed else 0) * m_loc, m_loc, SEED)
    A = torch.empty((n, m_loc), device="cuda").t()
    R = torch.zeros((n, n), device="cuda").t()

    launches_per_step = 0
    ref = None
    qr = None
    if args.impl == "b200":
        from later_b200 import qr
        from later_b200.tsqr import tsqr_rgsqrf
        if distributed:
            # The library's NCCL collectives get a non-blocking stream of their own: on the legacy default
            # stream they would synchronise implicitly with every blocking stream of the process.
            work_stream = torch.cuda.Stream()
            torch.cuda.set_stream(work_stream)
        ctx_main, ctx_stack = qr.Context(), qr.Context()
        if distributed:
            qr.comm_init(ctx_main)                 # the library's own NCCL communicator over the ranks

        def step():
            nonlocal launches_per_step
            if distributed:
                qr.later_rgsqrf_dist(ctx_main, m_loc, n, A, m_loc, R, n)
                launches_per_step = ctx_main.last_launch_count
            else:
                qr.later_rgsqrf(ctx_main, m_loc, n, A, m_loc, R, n)
                launches_per_step = ctx_main.last_launch_count
    else:
        try:
            ref = ReferenceLib()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({"impl": "reference", "unavailable": str(e)}))
            return 0
        work = torch.zeros(max(m_loc // 256 * 32 * n, 1 << 20) + (1 << 20), device="cuda")
        hwork = torch.zeros(m_loc * n, device="cuda", dtype=torch.float16)

        def step():
            ref.rgsqrf(m_loc, n, A, R, work, hwork)

    # ---- warm-up (graph capture, cuBLAS init for the reference, clocks)
    for _ in range(args.warmup):
        A.copy_(A0)
        step()
    barrier()

    # ---- timed region: EXACTLY K steps, device-timed with CUDA events around each factorisation
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for e0, e1 in evs:
        A.copy_(A0)                 # restore the in-place input (outside the event pair)
        if distributed:
            dist.barrier()
        e0.record()
        step()
        e1.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else {}
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = torch.tensor([sum(step_ms)], device="cuda", dtype=torch.float64)
    if distributed:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)      # max over ranks
    ms_per_step = float(total_ms.item()) / args.steps
    value = flops / (ms_per_step * 1e-3) / 1e12

    # ---- accuracy of the last step, GLOBAL (all-reduced over the row shards): reported, and a guard
    # against timing a broken run
    res2, nrm2, G = chunked_metrics(torch, A0, A, R)
    if distributed:
        dist.all_reduce(res2); dist.all_reduce(nrm2); dist.all_reduce(G)
    G.diagonal().sub_(1.0)
    back = float(torch.sqrt(res2 / nrm2))
    orth = float(torch.linalg.norm(G)) / n
    del G

    line = {
        "metric": "RGSQRF TFLOPS (2mn^2-2/3n^3)/s", "value": value, "unit": "TFLOPS",
        "n_gpus": N, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp32 in/out, fp16 tensor-core products, fp32 accumulate",
        "data": f"synthetic N(0,1): counter-based generator keyed by the global element index, seed {SEED} "
                "(a row shard equals those rows of the single-GPU matrix)",
        "config": {"workload": workload, "m": m, "n": n, "parallelism": f"dp{N}" if N > 1 else "single",
                   "l2": f"inputs ({m_loc * n * 4 / 2 ** 20:.0f} MiB per GPU) exceed the 126 MB L2; no flush needed",
                   "timing": "CUDA events around each factorisation, max over ranks; input restore outside"},
        "step_ms": step_ms, "wall_ms_per_step_incl_restore": t_wall * 1e3 / args.steps,
        "executed_tflops": 2.0 * m * n * n / (ms_per_step * 1e-3) / 1e12,
        "clocks": clocks,
        "backward_error": back, "orthogonality": orth,
        "accuracy_scope": "global matrix (all-reduced over ranks), last timed step, fp64 evaluation",
    }
    if args.impl == "reference":
        line["impl"] = "reference"
        line["reference_impl"] = "oracle/_ref/libref_later.so: unmodified reference CUDA sources, sm_100, cuBLAS 12.9"
        line["gpu_launches"] = None
        line["e2e"] = {"value": value, "unit": "TFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        if N == 1 and not (args.m and args.n) and not args.no_configs:
            del A0, A, R, work, hwork
            torch.cuda.empty_cache()
            line["configs"] = run_configs(torch, "reference", None, None, ref, peaks, peak_src, False)
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = lapack_baseline()
        print(json.dumps(line))
        return 0

    line["gpu_launches"] = launches_per_step * args.steps

    if distributed:
        # ---- per-GPU roofline of the sharded step (HBM-bound: SURVEY.md par.8d)
        gbs = algorithmic_bytes(m_loc, n) / (ms_per_step * 1e-3) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": "whole sharded step, per GPU (row-sharded recursion incl. its all-reduces)",
                            "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
                            "traffic": None, "bytes_per_gpu": algorithmic_bytes(m_loc, n),
                            "peak_source": f"{peak_src} MEASURED_PEAKS.json hbm_gbs"}
        # ---- the SAME global matrix on ONE GPU (rank 0, in-run): the anchor of the scaling claim, and
        # the single-GPU R to compare with
        anchor = torch.zeros(3, device="cuda", dtype=torch.float64)
        if rank == 0:
            T0 = keyed_normal(torch, m, n, 0, m, SEED)
            T = torch.empty((n, m), device="cuda").t()
            Rt = torch.zeros((n, n), device="cuda").t()
            c1 = qr.Context()
            t1 = timed(torch, lambda: qr.later_rgsqrf(c1, m, n, T, m, Rt, n), lambda: T.copy_(T0))
            anchor[0] = t1
            anchor[1] = float((R - Rt).abs().max() / Rt.abs().max())
            r2, n2, G1 = chunked_metrics(torch, T0, T, Rt)
            anchor[2] = float(torch.sqrt(r2 / n2))
            c1.close()
            del T0, T, Rt, G1
            torch.cuda.empty_cache()
        dist.broadcast(anchor, 0)
        # ---- the classical TSQR variant on the same shards, for comparison (fp16 back-multiplication)
        ctx_t = qr.Context()
        T = torch.empty((n, m_loc), device="cuda").t()
        Rt = torch.zeros((n, n), device="cuda").t()
        ts = []
        for i in range(6):
            T.copy_(A0)
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); tsqr_rgsqrf(m_loc, n, T, m_loc, Rt, n, ctxs=(ctx_t, ctx_stack)); e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        t_tsqr = torch.tensor([sorted(ts)[1]], device="cuda", dtype=torch.float64)
        dist.all_reduce(t_tsqr, op=dist.ReduceOp.MAX)
        r2, n2, Gt = chunked_metrics(torch, A0, T, Rt)
        dist.all_reduce(r2); dist.all_reduce(n2); dist.all_reduce(Gt)
        Gt.diagonal().sub_(1.0)
        line["tsqr_variant"] = {"ms_per_step": float(t_tsqr.item()), "backward_error": float(torch.sqrt(r2 / n2)),
                                "orthogonality": float(torch.linalg.norm(Gt)) / n,
                                "what": "local QR + all-gather of R + redundant stack QR + fp16 back-multiplication"}
        del T, Rt, Gt
        ctx_t.close()
        line["one_gpu_same_matrix"] = {"ms_per_step": float(anchor[0]), "backward_error": float(anchor[2]),
                                       "max_rel_diff_R_vs_sharded": float(anchor[1])}
        line["speedup_vs_1gpu"] = float(anchor[0]) / ms_per_step

    # ---- roofline of the dominant tensor kernel (top-level trailing-update GEMMs), timed live
    if rank == 0 and not distributed:
        h = n // 2
        Qh = torch.empty((n, m), device="cuda", dtype=torch.float16).normal_().t()
        Cg = torch.empty((h, h), device="cuda").t()
        Bh = torch.empty((h, h), device="cuda", dtype=torch.float16).normal_().t()

        def time_kernel(fn, reps=5):
            for _ in range(2):
                fn()
            ts = []
            for _ in range(reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            return sum(ts) / len(ts)
        t_gram = time_kernel(lambda: qr.gemm_gram(ctx_main, Qh, 0, h, h, h, Cg, None, 1))
        t_upd = time_kernel(lambda: qr.gemm_update(ctx_main, Qh, 0, h, Bh, A[:, h:], None, True))
        fl = 2.0 * h * h * m
        traffic, traffic_src = None, None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists():
            try:
                tjd = json.loads(tj.read_text())
                traffic = tjd.get("tc_gemm_gram_top_bytes")
                traffic_src = tjd.get("source")
            except Exception:
                traffic = None
        ach = fl / (t_gram * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "tc_gemm_kernel<256,gram> R12=Q1^T A2 top level",
                            "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": ach / peaks["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_src,
                            "peak_source": f"{peak_src} MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)",
                            "flops_per_launch": fl, "ms": t_gram}
        ach_u = fl / (t_upd * 1e-3) / 1e12
        line["roofline_update"] = {"bound": "tensor", "kernel": "tc_update_kernel<256> A2-=Q1 R12 top level",
                                   "achieved": ach_u, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                   "frac": ach_u / peaks["bf16_tflops"], "flops_per_launch": fl, "ms": t_upd}
        # the whole step against the SUSTAINED peak (a long step, not a kernel timed alone): the two
        # top-level GEMMs above are ~18 % of the step; panels and small recursion nodes are the rest
        line["roofline_step"] = {"bound": "tensor", "credited_tflops": value,
                                 "executed_tflops": line["executed_tflops"],
                                 "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                                 "frac_credited": value / peaks["bf16_tflops_sustained"],
                                 "frac_executed": line["executed_tflops"] / peaks["bf16_tflops_sustained"],
                                 "kernel_shares": "profiles/ (ncu launch list of this command, per round)"}
        line["whole_step_frac_of_tensor_peak"] = value / peaks["bf16_tflops_sustained"]
        del Qh, Cg, Bh

    # ---- end to end through the public host-buffer API (H2D + factorise + D2H every step)
    if not args.no_e2e:
        hA0 = torch.empty((n, m_loc), dtype=torch.float32).pin_memory()
        hA0.copy_(A0.t())
        hA = torch.empty((n, m_loc), dtype=torch.float32).pin_memory()
        hR = torch.zeros((n, n), dtype=torch.float32).pin_memory()
        e2e_steps = min(args.steps, 3)
        e2e_warm = 3                              # direct launch, graph capture, first replay
        times = []
        for i in range(e2e_steps + e2e_warm):
            hA.copy_(hA0)
            barrier()
            t0 = time.perf_counter()
            if distributed:
                A.t().copy_(hA0, non_blocking=True)
                qr.later_rgsqrf_dist(ctx_main, m_loc, n, A, m_loc, R, n)
                hA.copy_(A.t(), non_blocking=True)
                hR.copy_(R.t(), non_blocking=True)
                barrier()
            else:
                qr.later_rgsqrf_host(ctx_main, m_loc, n, hA.t(), m_loc, hR.t(), n)
            if i >= e2e_warm:
                times.append(time.perf_counter() - t0)
        t_e2e = torch.tensor([sum(times) / len(times)], device="cuda", dtype=torch.float64)
        if distributed:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
            r_back = n * n
        else:
            # the host entry point sends back the block upper triangle of R (include/later_b200.h)
            c = max(min(n, 256), n // 16)
            r_back = c * c * (n // c) * (n // c + 1) // 2
        line["e2e"] = {"value": flops / float(t_e2e.item()) / 1e12, "unit": "TFLOPS",
                       "h2d_bytes_per_step": 4 * m_loc * n * shards,
                       "d2h_bytes_per_step": 4 * (m_loc * n + r_back) * shards,
                       "ms_per_step": float(t_e2e.item()) * 1e3,
                       "step_ms": [t * 1e3 for t in times],
                       "api": ("H2D of the row block + later_rgsqrf_dist + D2H of Q and R (pinned buffers)" if distributed else
                               "later_rgsqrf_host (pinned host A in, Q and R out)")}
        if distributed:
            # what the box's host side gives when every rank copies at once and nothing computes: the
            # ceiling of the end-to-end number above
            probe = []
            for direction in ("h2d", "d2h"):
                barrier()
                t0 = time.perf_counter()
                if direction == "h2d":
                    A.t().copy_(hA0, non_blocking=True)
                else:
                    hA.copy_(A.t(), non_blocking=True)
                torch.cuda.synchronize()
                dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
                probe.append(4.0 * m_loc * n / float(dt.item()) / 1e9)
            line["e2e"]["pcie_probe"] = {"h2d_gbs_per_rank_all_ranks_copying": probe[0],
                                         "d2h_gbs_per_rank_all_ranks_copying": probe[1],
                                         "aggregate_h2d_gbs": probe[0] * shards, "aggregate_d2h_gbs": probe[1] * shards,
                                         "copy_only_floor_ms": (4.0 * m_loc * n / 1e9) * (1 / probe[0] + 1 / probe[1]) * 1e3}
        if not distributed:
            # same kernels on the same input: the host path must reproduce the device path bit for bit
            A.copy_(A0)
            qr.later_rgsqrf(ctx_stack, m_loc, n, A, m_loc, R, n)
            torch.cuda.synchronize()
            same_q = torch.equal(hA.cuda(), A.t())                       # every column of Q
            c = max(min(n, 256), n // 16)
            blk = torch.arange(n, device="cuda") // c
            mask = blk[:, None] <= blk[None, :]                          # the block upper triangle sent back
            same_r = torch.equal(hR.cuda().t()[mask], R[mask])
            line["e2e"]["matches_device_path"] = bool(same_q and same_r)
            line["e2e"]["matches_device_path_scope"] = "all of Q, all transferred blocks of R, bit for bit"
        del hA0, hA, hR

    # ---- the other single-GPU configurations of BASELINE.json, same run
    if rank == 0 and not distributed and N == 1 and not (args.m and args.n) and not args.no_configs:
        del A0, A, R
        torch.cuda.empty_cache()
        line["configs"] = run_configs(torch, "b200", qr, ctx_stack, None, peaks, peak_src, False)
        c4 = line["configs"].get("C4_1048576x1024_one_gpu")
        if c4:      # (kept under its round-1 name too: the anchor of the N > 1 lines)
            line["tall_skinny_1gpu"] = {"workload": "rgsqrf_1048576x1024 on one GPU", "ms_per_step": c4["ms"],
                                        "value": c4["tflops"], "unit": "TFLOPS",
                                        "algorithmic_gbs": c4["algorithmic_gbs"],
                                        "hbm_roofline_frac": c4["hbm_roofline_frac"],
                                        "peak_source": f"{peak_src} MEASURED_PEAKS.json hbm_gbs"}
    if rank == 0 and not args.no_cpu_baseline and not distributed:
        line["cpu_baseline"] = lapack_baseline()
    if rank == 0:
        print(json.dumps(line))
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
