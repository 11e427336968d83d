#!/usr/bin/env python
"""One Always-mode offload call (after one warm-up) -- use with B200_ENGINE_TRACE=1."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
kind, d = sys.argv[1], int(sys.argv[2])
dt = "f64" if kind[0] == "d" else "f32"; s = 8 if dt == "f64" else 4
ctx = blob.Context(0)
ops = [ctx.alloc(0, d * d * s) for _ in range(3)]
for (h, _), v in zip(ops, (1.5, 0.5, 0.0)): blob.host_view(h, d * d, dt)[:] = v
(hA, dA), (hB, dB), (hC, dC) = ops
for i in range(2):
    t0 = time.perf_counter()
    ctx.gemm_offload(dt, d, d, d, 1.0, hA, dA, d, hB, dB, d, 0.0, hC, dC, d)
    print(f"call {i}: {(time.perf_counter()-t0)*1e3:.2f} ms  C[0]={blob.host_view(hC, 4, dt)[0]} (want {1.5*0.5*d})", file=sys.stderr)
