#!/usr/bin/env python
"""Median wall time of the Always-mode offload call (host buffers -> result on the host).
usage: e2e_time.py <dgemm|sgemm> <dim> [reps]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
kind, d = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 7
dt = "f64" if kind[0] == "d" else "f32"; s = 8 if dt == "f64" else 4
ctx = blob.Context(0)
ops = [ctx.alloc(0, d * d * s) for _ in range(3)]
for (h, _), v in zip(ops, (1.5, 0.5, 0.0)): blob.host_view(h, d * d, dt)[:] = v
(hA, dA), (hB, dB), (hC, dC) = ops
if dt == "f32": ctx.sgemm_prepare(d, d, d, d, d)
ts = []
for i in range(reps + 1):
    t0 = time.perf_counter()
    ctx.gemm_offload(dt, d, d, d, 1.0, hA, dA, d, hB, dB, d, 0.0, hC, dC, d)
    ts.append(time.perf_counter() - t0)
ts = sorted(ts[1:]); med = ts[len(ts) // 2]
ok = abs(float(blob.host_view(hC, d * d, dt)[d * d - 1]) - 1.5 * 0.5 * d) <= 1e-3 * d
print(f"{kind} {d}^3 slabs_env={os.environ.get('B200_ENGINE_SLABS', 'default')}: median {med*1e3:.2f} ms  min {ts[0]*1e3:.2f}  {(2.0*d**3+d*d)/med*1e-12:.2f} TFLOP/s  result ok={ok}")
