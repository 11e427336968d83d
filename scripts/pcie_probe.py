#!/usr/bin/env python
"""Pinned H2D / D2H bandwidth through the library's copy lanes, alone and concurrently."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
ctx = blob.Context(0)
n = 512 << 20
h1, d1 = ctx.alloc(1, n); h2, d2 = ctx.alloc(1, n)
blob.host_view(h1, n // 8, "f64")[:] = 1.0; blob.host_view(h2, n // 8, "f64")[:] = 2.0
def t(f, reps=3):
    f(); ctx.sync(); t0 = time.perf_counter()
    for _ in range(reps): f()
    ctx.sync(); return (time.perf_counter() - t0) / reps
a = t(lambda: ctx.h2d(d1, h1, n, 0)); print(f"H2D 512MiB: {n/a/1e9:.1f} GB/s")
b = t(lambda: ctx.d2h(h2, d2, n, 2)); print(f"D2H 512MiB: {n/b/1e9:.1f} GB/s")
c = t(lambda: (ctx.h2d(d1, h1, n, 0), ctx.d2h(h2, d2, n, 2))); print(f"H2D+D2H concurrent: {2*n/c/1e9:.1f} GB/s total ({c*1e3:.1f} ms)")
for piece in (64 << 20, 16 << 20, 4 << 20, 1 << 20):
    k = n // piece
    e = t(lambda: [ctx.h2d(d1 + i * piece, h1 + i * piece, piece, 0) for i in range(k)])
    print(f"H2D in {piece>>20} MiB pieces: {n/e/1e9:.1f} GB/s")
