"""torchrun worker: row-sharded TSQR on real GPUs (NCCL), checked against the single-GPU result.
Usage: torchrun --nproc-per-node P tests/helpers/tsqr_worker.py m n"""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from later_b200 import qr  # noqa: E402
from later_b200.tsqr import tsqr_rgsqrf  # noqa: E402


def main():
    m, n = int(sys.argv[1]), int(sys.argv[2])
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(1234)          # same global matrix on every rank
    A_glob = torch.randn(m, n, device="cuda", generator=g)
    mloc = m // world
    A0 = A_glob[rank * mloc:(rank + 1) * mloc].clone()
    A = qr.to_colmajor(A0)
    R = qr.colmajor_empty(n, n)
    ctxs = (qr.Context(), qr.Context())
    tsqr_rgsqrf(mloc, n, A, mloc, R, n, ctxs=ctxs)
    torch.cuda.synchronize()
    # every rank must hold the same R, bit for bit
    Rs = [torch.empty_like(R.contiguous()) for _ in range(world)]
    dist.all_gather(Rs, R.contiguous())
    same = all(torch.equal(Rs[0], x) for x in Rs)
    # global checks: ||A - Q R|| and ||I - Q^T Q|| assembled with all-reduces of small quantities
    res2 = torch.linalg.norm((A0 - A @ R).double()) ** 2
    nrm2 = torch.linalg.norm(A0.double()) ** 2
    G = (A.t() @ A).double()
    dist.all_reduce(res2); dist.all_reduce(nrm2); dist.all_reduce(G)
    back = float(torch.sqrt(res2 / nrm2))
    G.diagonal().sub_(1.0)
    orth = float(torch.linalg.norm(G) / n)
    ok = same and back < 5e-4 and orth < 5e-5 and bool((R.diagonal() > 0).all()) \
        and float(torch.tril(R, -1).abs().max()) == 0.0
    # same factorisation with the row block arriving from pinned host memory during the local QR
    hA = torch.empty((n, mloc), dtype=torch.float32).pin_memory()
    hA.copy_(A0.t())
    A2 = qr.colmajor_empty(mloc, n)
    R2 = qr.colmajor_empty(n, n)
    for _ in range(3):      # direct launch, graph capture, graph replay
        A2.fill_(float("nan"))
        tsqr_rgsqrf(mloc, n, A2, mloc, R2, n, ctxs=ctxs, host_A=hA.t())
        torch.cuda.synchronize()
        ok = ok and torch.equal(A2, A) and torch.equal(R2, R)
    # the row-sharded recursion (all-reduces inside): single-GPU accuracy, identical R on every rank,
    # direct / capturing / replaying call bit-identical
    # (own non-blocking stream: the library's NCCL collectives must not sit on the legacy default stream,
    # which synchronises implicitly with every blocking stream of the process)
    side = torch.cuda.Stream()
    dctx = qr.Context(stream=side)
    qr.comm_init(dctx)
    A3 = qr.colmajor_empty(mloc, n)
    R3 = qr.colmajor_empty(n, n)
    outs = []
    for _ in range(3):
        A3.copy_(A0)
        R3.fill_(float("nan"))
        side.wait_stream(torch.cuda.current_stream())
        qr.later_rgsqrf_dist(dctx, mloc, n, A3, mloc, R3, n)
        side.synchronize()
        outs.append((A3.clone(), R3.clone()))
    ok = ok and all(torch.equal(outs[0][0], a) and torch.equal(outs[0][1], r) for a, r in outs[1:])
    Rs3 = [torch.empty_like(R3.contiguous()) for _ in range(world)]
    dist.all_gather(Rs3, R3.contiguous())
    same3 = all(torch.equal(Rs3[0], x) for x in Rs3)
    res3 = torch.linalg.norm((A0 - A3 @ R3).double()) ** 2
    G3 = (A3.t() @ A3).double()
    dist.all_reduce(res3); dist.all_reduce(G3)
    back3 = float(torch.sqrt(res3 / nrm2))
    G3.diagonal().sub_(1.0)
    orth3 = float(torch.linalg.norm(G3) / n)
    ok = ok and same3 and bool((R3.diagonal() > 0).all()) and float(torch.tril(R3, -1).abs().max()) == 0.0
    if rank == 0:
        # against the single-GPU factorisation of the whole matrix
        c = qr.Context()
        A1 = qr.to_colmajor(A_glob)
        R1 = qr.colmajor_empty(n, n)
        qr.later_rgsqrf(c, m, n, A1, m, R1, n)
        rdiff = float((R - R1).abs().max() / R1.abs().max())
        ok = ok and rdiff < 5e-3
        # the row-sharded recursion must be as accurate as the single-GPU factorisation (within 2x)
        back1 = float(torch.linalg.norm((A_glob - A1 @ R1).double()) / torch.linalg.norm(A_glob.double()))
        G1 = (A1.t() @ A1).double()
        G1.diagonal().sub_(1.0)
        orth1 = float(torch.linalg.norm(G1) / n)
        rdiff3 = float((R3 - R1).abs().max() / R1.abs().max())
        ok = ok and back3 <= 2 * back1 + 1e-7 and orth3 <= 2 * orth1 + 1e-8 and rdiff3 < 5e-3
        print(f"DIST world={world} {m}x{n}: same_R={same3} backward={back3:.3e} (1 GPU {back1:.3e}) "
              f"orth/n={orth3:.3e} (1 GPU {orth1:.3e}) |R-R1|/|R1|={rdiff3:.3e}", flush=True)
        print(f"TSQR world={world} {m}x{n}: same_R={same} backward={back:.3e} orth/n={orth:.3e} "
              f"|R-R1|/|R1|={rdiff:.3e} {'OK' if ok else 'FAIL'}", flush=True)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
