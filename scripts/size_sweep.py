#!/usr/bin/env python
"""Device-resident per-call time across sizes (CUDA events, 20 calls after 5 warm-ups).
usage: size_sweep.py <dgemm|sgemm|dgemv|sgemv> size [size ...]   (square shapes)"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
kind = sys.argv[1]
dt = "f64" if kind[0] == "d" else "f32"
tdt = torch.float64 if dt == "f64" else torch.float32
ctx = blob.Context(0)
out = []
for d in map(int, sys.argv[2:]):
    a = torch.rand(d * d, dtype=tdt, device="cuda")
    if kind.endswith("gemm"):
        b, c = torch.rand(d * d, dtype=tdt, device="cuda"), torch.zeros(d * d, dtype=tdt, device="cuda")
        call = lambda: ctx.gemm(dt, d, d, d, 1.0, a, d, b, d, 0.0, c, d)
        flops = 2 * d ** 3 + d * d
    else:
        b, c = torch.rand(d, dtype=tdt, device="cuda"), torch.zeros(d, dtype=tdt, device="cuda")
        call = lambda: ctx.gemv(dt, d, d, 1.0, a, d, b, 1, 0.0, c, 1)
        flops = 2 * d * d + d
    torch.cuda.synchronize()
    for _ in range(5): call()
    ctx.sync(); ctx.timer_start()
    for _ in range(20): call()
    us = ctx.timer_stop() / 20 * 1e3
    out.append(f"{d}:{us:.1f}us/{flops/us*1e-3:.0f}GF[{ctx.last_kernel().split('_')[-2] if 'dmma' in ctx.last_kernel() else ctx.last_kernel()[-14:]}]")
print(kind, os.environ.get("B200_DGEMM_TILE", ""), os.environ.get("B200_SGEMM", ""), " ".join(out))
