#!/usr/bin/env python
"""Where do the Unified-mode outliers come from?  Repeats the exact choreography of
gpu::gemm_gpu<T> in Unified mode (B200/gemm.hh: acquire managed blocks, CPU fill,
prefetch to the GPU, 10 kernels, prefetch C back, sync, release) and times every phase
with the host clock, syncing between phases.

usage: unified_phases.py [dims ...]   (default 1024 2048 4096)  env REPS (default 12)
"""
import importlib
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
import numpy as np  # noqa: E402

dims = [int(a) for a in sys.argv[1:]] or [1024, 2048, 4096]
reps = int(os.environ.get("REPS", "12"))
one_lane = os.environ.get("PF_LANES", "3") == "1"   # all three prefetches on one lane instead of three
kinds = os.environ.get("KINDS", "gemm,gemv").split(",")
ctx = blob.Context(0)
U = blob.OFFLOAD_UNIFIED
for kind in kinds:
    for d in dims:
        cnt = (d * d, d * d, d * d) if kind == "gemm" else (d * d, d, d)
        rows = []
        for rep in range(reps):
            t = [time.perf_counter()]
            ops = [ctx.alloc(U, c * 8) for c in cnt]
            t.append(time.perf_counter())  # alloc
            views = [blob.host_view(h, c, "f64") for (h, _), c in zip(ops, cnt)]
            views[0][:] = 1.0 / 7.0
            views[1][:] = 1.0 / 3.0
            views[2][:] = 0.0
            t.append(time.perf_counter())  # cpu fill
            for lane, ((h, _), c) in enumerate(zip(ops, cnt)):
                ctx.prefetch(h, c * 8, True, 0 if one_lane else lane)
            ctx.sync()
            t.append(time.perf_counter())  # prefetch to gpu
            for _ in range(10):
                if kind == "gemm":
                    ctx.gemm("f64", d, d, d, 1.0, ops[0][0], d, ops[1][0], d, 0.0, ops[2][0], d)
                else:
                    ctx.gemv("f64", d, d, 1.0, ops[0][0], d, ops[1][0], 1, 0.0, ops[2][0], 1)
            ctx.sync()
            t.append(time.perf_counter())  # kernels
            ctx.prefetch(ops[2][0], cnt[2] * 8, False, 2)
            ctx.sync()
            t.append(time.perf_counter())  # prefetch back
            chk = float(views[2][0]) + float(views[2][-1])
            t.append(time.perf_counter())  # cpu read
            for h, dv in ops:
                ctx.free(U, h, dv)
            t.append(time.perf_counter())  # release
            rows.append([(b - a) * 1e3 for a, b in zip(t, t[1:])])
            del views
        a = np.array(rows)
        names = ["alloc", "cpu_fill", "pf_to_gpu", "10_kernels", "pf_C_back", "cpu_read", "release"]
        print(f"--- PF_LANES={1 if one_lane else 3} d{kind} {d}: ms per phase over {reps} reps (median | max | rep-of-max)  checksum {chk:.3f}")
        for j, nm in enumerate(names):
            print(f"   {nm:11s} {np.median(a[:, j]):9.3f} | {a[:, j].max():9.3f} | {int(a[:, j].argmax())}")
        timed = a[:, 2] + a[:, 3] + a[:, 4]
        print(f"   timed(sum)  {np.median(timed):9.3f} | {timed.max():9.3f} | all: " + " ".join(f"{x:.1f}" for x in timed))
ctx.close()
