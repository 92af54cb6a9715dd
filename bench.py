#!/usr/bin/env python
"""RGSQRF benchmark (BASELINE.json's metric): TFLOPS = (2 m n^2 - 2/3 n^3) / t, the formula the
reference prints (reference test/test_qr.cu:85-86).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

N = 1 : 16384 x 16384 on one B200 (BASELINE.json configs[1]).
N > 1 : 1048576 x 1024, row-sharded over N GPUs, TSQR combine (configs[3]); launched by
        torch.distributed.run, one rank per GPU, NCCL.  Strong scaling (total work fixed).
One JSON line on stdout (rank 0).  Inputs are synthetic N(0,1), resident in HBM before the timed
region; the matrices (1-4 GiB) are larger than the 126 MB L2, so no explicit L2 flush is needed.

--impl reference times the UNMODIFIED reference (oracle/_ref/libref_later.so: the reference's CUDA
sources compiled for sm_100 against cuBLAS 12.9 by oracle/Makefile) on the same workload.  The
reference has no CPU path; host LAPACK sgeqrf+sorgqr is the CPU stand-in BASELINE.json names and is
reported as `cpu_baseline` on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
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
            "sample": f"host LAPACK sgeqrf+sorgqr (scipy OpenBLAS) on {n_sample}x{n_sample} N(0,1), "
                      f"{dt:.1f} s; stands in for a CPU path the reference does not have"}


def measured_peaks() -> tuple[dict, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


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

    def rgsqrf(self, m, n, A, R, work, hwork):
        rc = self.lib.ref_later_rgsqrf(m, n, A.data_ptr(), m, R.data_ptr(), n, work.data_ptr(),
                                       work.numel(), hwork.data_ptr(), hwork.numel())
        if rc != 0:
            raise RuntimeError(f"reference later_rgsqrf: cuda error {rc}")


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
        workload = f"rgsqrf_1048576x1024_rowsharded_tsqr_dp{N}"
    if args.m and args.n:
        m, n = args.m, args.n
        workload = f"rgsqrf_{m}x{n}_override"
    shards = world if distributed else 1
    m_loc = m // shards
    flops = credited_flops(m, n)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    gen = torch.Generator(device="cuda").manual_seed(3000 + (rank if distributed else 0))
    A0 = torch.empty((n, m_loc), device="cuda").normal_(generator=gen).t()     # column-major m_loc x n
    A = torch.empty((n, m_loc), device="cuda").t()
    R = torch.zeros((n, n), device="cuda").t()

    launches_per_step = 0
    if args.impl == "b200":
        from later_b200 import qr
        from later_b200.tsqr import tsqr_rgsqrf
        ctx_main, ctx_stack = qr.Context(), qr.Context()

        def step():
            nonlocal launches_per_step
            if distributed:
                tsqr_rgsqrf(m_loc, n, A, m_loc, R, n, ctxs=(ctx_main, ctx_stack))
                launches_per_step = 2 * ctx_main.last_launch_count + ctx_stack.last_launch_count
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

    # ---- accuracy of the last step (reported, and a guard against timing a broken run)
    if distributed:
        back = orth = None
    else:
        back = float(torch.linalg.norm((A0 - A @ R).double()) / torch.linalg.norm(A0.double())) if m * n <= 2 ** 29 else None
        orth = None
    line = {
        "metric": "RGSQRF TFLOPS (2mn^2-2/3n^3)/s", "value": value, "unit": "TFLOPS",
        "n_gpus": N, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp32 in/out, fp16 tensor-core products, fp32 accumulate",
        "data": "synthetic N(0,1), seed 3000",
        "config": {"workload": workload, "m": m, "n": n, "parallelism": f"dp{N}" if N > 1 else "single",
                   "l2": "inputs (>= 1 GiB) exceed the 126 MB L2; no flush needed",
                   "timing": "CUDA events around each factorisation, max over ranks; input restore outside"},
        "step_ms": step_ms, "wall_ms_per_step_incl_restore": t_wall * 1e3 / args.steps,
        "executed_tflops": 2.0 * m * n * n / (ms_per_step * 1e-3) / 1e12,
        "clocks": clocks,
    }
    if back is not None:
        line["backward_error"] = back
    if args.impl == "reference":
        line["impl"] = "reference"
        line["reference_impl"] = "oracle/_ref/libref_later.so: unmodified reference CUDA sources, sm_100, cuBLAS 12.9"
        line["gpu_launches"] = None
        line["e2e"] = {"value": value, "unit": "TFLOPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = lapack_baseline()
        print(json.dumps(line))
        return 0

    line["gpu_launches"] = launches_per_step * args.steps

    # ---- roofline of the dominant kernel (top-level trailing-update GEMMs), timed live
    if rank == 0 and not distributed:
        peaks, src = measured_peaks()
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
        traffic = None
        tj = ROOT / "profiles" / "traffic.json"
        if tj.exists():
            try:
                traffic = json.loads(tj.read_text()).get("tc_gemm_gram_top_bytes")
            except Exception:
                traffic = None
        ach = fl / (t_gram * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "tc_gemm_kernel<256,gram> R12=Q1^T A2 top level",
                            "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": ach / peaks["bf16_tflops"], "traffic": traffic,
                            "peak_source": f"{src} MEASURED_PEAKS.json bf16_tflops (burst: kernel timed alone)",
                            "flops_per_launch": fl, "ms": t_gram}
        ach_u = fl / (t_upd * 1e-3) / 1e12
        line["roofline_update"] = {"bound": "tensor", "kernel": "tc_gemm_kernel<256,update> A2-=Q1 R12 top level",
                                   "achieved": ach_u, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                                   "frac": ach_u / peaks["bf16_tflops"], "flops_per_launch": fl, "ms": t_upd}
        line["whole_step_frac_of_tensor_peak"] = value / peaks["bf16_tflops_sustained"]
        del Qh, Cg, Bh

    # ---- the tall-skinny workload on ONE GPU, so that the N>1 lines have their 1-GPU anchor
    if rank == 0 and not distributed and N == 1 and not (args.m and args.n):
        mt, nt = 1048576, 1024
        T0 = torch.empty((nt, mt), device="cuda").normal_(generator=gen).t()
        T = torch.empty((nt, mt), device="cuda").t()
        Rt = torch.zeros((nt, nt), device="cuda").t()
        ts = []
        for i in range(8):
            T.copy_(T0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); qr.later_rgsqrf(ctx_stack, mt, nt, T, mt, Rt, nt); e1.record()
            torch.cuda.synchronize()
            if i >= 3:                       # direct launch, graph capture, first replay
                ts.append(e0.elapsed_time(e1))
        t_ts = sorted(ts)[len(ts) // 2]     # median of 5
        peaks, src = measured_peaks()
        line["tall_skinny_1gpu"] = {
            "workload": "rgsqrf_1048576x1024 on one GPU", "ms_per_step": t_ts,
            "value": credited_flops(mt, nt) / (t_ts * 1e-3) / 1e12, "unit": "TFLOPS",
            "algorithmic_gbs": (8.0 * mt * nt + 4.0 * nt * nt) / (t_ts * 1e-3) / 1e9,
            "hbm_roofline_frac": (8.0 * mt * nt + 4.0 * nt * nt) / (t_ts * 1e-3) / 1e9 / peaks["hbm_gbs"],
            "peak_source": f"{src} MEASURED_PEAKS.json hbm_gbs"}
        del T0, T, Rt

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
                # the row block crosses PCIe while the local factorisation already runs on the
                # columns that have arrived; Q leaves after the TSQR back-multiplication
                tsqr_rgsqrf(m_loc, n, A, m_loc, R, n, ctxs=(ctx_main, ctx_stack), host_A=hA0.t())
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
            c = max(128, n // 16)
            r_back = c * c * (n // c) * (n // c + 1) // 2
        line["e2e"] = {"value": flops / float(t_e2e.item()) / 1e12, "unit": "TFLOPS",
                       "h2d_bytes_per_step": 4 * m_loc * n * shards,
                       "d2h_bytes_per_step": 4 * (m_loc * n + r_back) * shards,
                       "ms_per_step": float(t_e2e.item()) * 1e3,
                       "step_ms": [t * 1e3 for t in times],
                       "api": ("tsqr_rgsqrf(host_A=pinned) + D2H of Q and R" if distributed else
                               "later_rgsqrf_host (pinned host A in, Q and R out)")}
        if not distributed:
            # same kernels on the same input: the host path must reproduce the device path bit for bit
            A.copy_(A0)
            qr.later_rgsqrf(ctx_stack, m_loc, n, A, m_loc, R, n)
            torch.cuda.synchronize()
            cols = slice(n - 256, n)
            same_q = torch.equal(hA[cols].cuda(), A.t()[cols]) and torch.equal(hA[:256].cuda(), A.t()[:256])
            same_r = torch.equal(torch.triu(hR.cuda().t()), torch.triu(R))
            line["e2e"]["matches_device_path"] = bool(same_q and same_r)
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
