"""Row-sharded recursion with a ONE-rank communicator: exercises NCCL inside the launch sequence
(direct, stream capture, graph replay) on a single GPU.  Always run under `timeout`."""
import ctypes as C
import sys
import time
import torch
sys.path.insert(0, '.')
from later_b200 import qr
from later_b200._lib import lib

m, n = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8192, 512)
ctx = qr.Context()
buf = (C.c_ubyte * 128)()
print("unique id rc", lib.later_b200_comm_unique_id(buf), flush=True)
print("comm_init rc", lib.later_b200_comm_init(ctx._h, 1, 0, bytes(buf)), flush=True)
g = torch.Generator(device="cuda").manual_seed(1)
A0 = torch.randn(m, n, device="cuda", generator=g)
ref = qr.Context()
A1 = qr.to_colmajor(A0); R1 = qr.colmajor_empty(n, n)
qr.later_rgsqrf(ref, m, n, A1, m, R1, n)
torch.cuda.synchronize()
for i in range(4):
    A = qr.to_colmajor(A0); R = qr.colmajor_empty(n, n)
    t0 = time.time()
    qr.later_rgsqrf_dist(ctx, m, n, A, m, R, n)
    print("call", i, "enqueued", f"{time.time() - t0:.3f}s", flush=True)
    torch.cuda.synchronize()
    print("call", i, "done: equals single-GPU path:", torch.equal(A, A1), torch.equal(R, R1),
          "launches", ctx.last_launch_count, "graph", ctx.graph_stats(), flush=True)
print("dist1 ok", flush=True)
