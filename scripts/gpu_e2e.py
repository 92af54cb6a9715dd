"""End-to-end host-buffer path: PCIe rates of the box, then later_rgsqrf_host wall times per call."""
import sys, time, torch
sys.path.insert(0, '.')
from later_b200 import qr

m = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 6

gib = 1 << 28   # floats
h0 = torch.empty(gib, dtype=torch.float32).pin_memory()
h1 = torch.empty(gib, dtype=torch.float32).pin_memory()
d0 = torch.empty(gib, device='cuda'); d1 = torch.empty(gib, device='cuda')
s0, s1 = torch.cuda.Stream(), torch.cuda.Stream()
def wall(fn):
    torch.cuda.synchronize(); t = time.perf_counter(); fn(); torch.cuda.synchronize()
    return time.perf_counter() - t
for _ in range(2):
    t_in = wall(lambda: d0.copy_(h0, non_blocking=True))
    t_out = wall(lambda: h1.copy_(d1, non_blocking=True))
    def both():
        with torch.cuda.stream(s0): d0.copy_(h0, non_blocking=True)
        with torch.cuda.stream(s1): h1.copy_(d1, non_blocking=True)
    t_both = wall(both)
print(f"PCIe 1 GiB: H2D {1.0737/t_in:.1f} GB/s, D2H {1.0737/t_out:.1f} GB/s, both at once {t_both*1e3:.1f} ms "
      f"({1.0737/t_both:.1f} GB/s each way)", flush=True)
del h0, h1, d0, d1

ctx = qr.Context()
hA0 = torch.empty((n, m), dtype=torch.float32).pin_memory(); hA0.uniform_()
hA = torch.empty((n, m), dtype=torch.float32).pin_memory()
hR = torch.zeros((n, n), dtype=torch.float32).pin_memory()
flops = 2.0 * m * n * n - 2.0 / 3.0 * n ** 3
for i in range(calls):
    hA.copy_(hA0)
    t = wall(lambda: qr.later_rgsqrf_host(ctx, m, n, hA.t(), m, hR.t(), n))
    print(f"call {i}: {t*1e3:8.2f} ms  {flops/t/1e12:7.1f} TFLOPS e2e  launches {ctx.last_launch_count}", flush=True)
A = hA0.cuda().t(); Q = hA.cuda().t(); R = torch.triu(hR.cuda().t())
print("backward", (torch.linalg.matrix_norm(Q @ R - A) / torch.linalg.matrix_norm(A)).item(),
      "orth/n", (torch.linalg.matrix_norm(Q.t() @ Q - torch.eye(n, device='cuda')) / n).item())
