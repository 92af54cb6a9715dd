"""torchrun worker: smallest possible exercise of later_rgsqrf_dist over 2+ ranks, with progress prints.
Always run under `timeout`."""
import os
import sys
import time
import torch
import torch.distributed as dist
sys.path.insert(0, '.')
from later_b200 import qr

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl", device_id=torch.device("cuda", torch.cuda.current_device()))
m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (4096, 256)
mloc = m // world


def say(*a):
    print(f"[rank {rank}]", *a, flush=True)


stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = qr.Context(stream=stream)
    say("context on stream", hex(stream.cuda_stream))
    qr.comm_init(ctx)
    say("comm_init done")
    g = torch.Generator(device="cuda").manual_seed(7)
    A_glob = torch.randn(m, n, device="cuda", generator=g)
    A0 = A_glob[rank * mloc:(rank + 1) * mloc].clone()
    A = qr.colmajor_empty(mloc, n)
    R = qr.colmajor_empty(n, n)
    for i in range(4):
        A.copy_(A0)
        t0 = time.time()
        qr.later_rgsqrf_dist(ctx, mloc, n, A, mloc, R, n)
        say("call", i, "enqueued", f"{time.time() - t0:.3f}s")
        stream.synchronize()
        say("call", i, "done", ctx.graph_stats())
    res2 = torch.linalg.norm((A0 - A @ R).double()) ** 2
    nrm2 = torch.linalg.norm(A0.double()) ** 2
    G = (A.t() @ A).double()
stream.synchronize()
dist.all_reduce(res2); dist.all_reduce(nrm2); dist.all_reduce(G)
G.diagonal().sub_(1.0)
say("backward", float(torch.sqrt(res2 / nrm2)), "orth", float(torch.linalg.norm(G) / n))
dist.barrier()
dist.destroy_process_group()
