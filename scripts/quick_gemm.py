#!/usr/bin/env python
"""Device-resident timing of one GEMM/GEMV shape through the C ABI (CUDA events).
usage: quick_gemm.py <dgemm|sgemm|dgemv|sgemv> M N [K] [iters]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
kind = sys.argv[1]
dt = "f64" if kind[0] == "d" else "f32"
tdt = torch.float64 if dt == "f64" else torch.float32
ctx = blob.Context(0)
def fill(n, div):
    return torch.randint(0, 100, (n,), device="cuda").to(tdt) / div
if kind.endswith("gemm"):
    m, n, k = map(int, sys.argv[2:5]); iters = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    a, b, c = fill(m * k, 7.0), fill(k * n, 3.0), torch.zeros(m * n, dtype=tdt, device="cuda")
    call = lambda: ctx.gemm(dt, m, n, k, 1.0, a, m, b, k, 0.0, c, m)
    flops = 2 * m * n * k + m * n; nbytes = (m * k + k * n + m * n) * a.element_size()
else:
    m, n = map(int, sys.argv[2:4]); iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    a, b, c = fill(m * n, 7.0), fill(n, 3.0), torch.zeros(m, dtype=tdt, device="cuda")
    call = lambda: ctx.gemv(dt, m, n, 1.0, a, m, b, 1, 0.0, c, 1)
    flops = 2 * m * n + m; nbytes = (m * n + n + m) * a.element_size()
torch.cuda.synchronize()
for _ in range(3): call()
ctx.sync(); ctx.timer_start()
for _ in range(iters): call()
ms = ctx.timer_stop() / iters
print(f"{kind} {sys.argv[2:]} kernel={ctx.last_kernel()} {ms:.4f} ms  {flops/ms*1e-6:.1f} GFLOP/s  {nbytes/ms*1e-6:.1f} GB/s")
