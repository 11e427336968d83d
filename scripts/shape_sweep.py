#!/usr/bin/env python
"""Device-resident kernel time for every problem type the reference sweeps
(include/doGemm.hh:63-291, include/doGemv.hh:63-181) at a given upper dimension D,
against the roofline that bounds each shape: min(pipe peak, AI x HBM bandwidth).

usage: shape_sweep.py [--gemm-dims 8192,2048] [--gemv-dims 32768,8192] [--md out.md]

Operands are rotated through enough copies to exceed the 126 MB L2, so a shape whose
working set fits in L2 is still measured against HBM (bench rule: flush or exceed L2).
Time = CUDA events on the library's compute stream around `iters` calls after 3 warm-ups.
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
L2_BYTES = 126 << 20


def gemm_types(D):
    s = max(1, D // 16)
    return [("square M=N=K", D, D, D), ("M=N, M=16K", D, D, s), ("M=N, K=32", D, D, 32),
            ("M=N, K=16M", s, s, D), ("M=N=32, K", 32, 32, D), ("M=16K, K=N", D, s, s),
            ("M, K=N=32", D, 32, 32), ("M=K, N=16K", s, D, s), ("M=K=32, N", 32, D, 32)]


def gemv_types(D):
    s = max(1, D // 16)
    return [("square M=N", D, D), ("M=16N", D, s), ("M, N=32", D, 32), ("N=16M", s, D), ("M=32, N", 32, D)]


def fill(n, div, tdt):
    return torch.randint(0, 100, (n,), device="cuda").to(tdt) / div


def time_calls(ctx, calls, iters):
    torch.cuda.synchronize()  # operands were produced on torch's stream
    for i in range(3):
        calls[i % len(calls)]()
    ctx.sync()
    ctx.timer_start()
    for i in range(iters):
        calls[i % len(calls)]()
    return ctx.timer_stop() / iters  # ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gemm-dims", default="8192,2048")
    ap.add_argument("--gemv-dims", default="32768,8192")
    ap.add_argument("--md", default="")
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    hbm = peaks["hbm_gbs"]
    ctx = blob.Context(0)
    pipe = {"f64": max(ctx.probe_peak("dmma"), ctx.probe_peak("dfma")),
            "f32_ffma": ctx.probe_peak("ffma"), "f32_tc": peaks["bf16_tflops"] * 1e3 / 6.0}
    rows = []
    for dt, tdt, es in (("f64", torch.float64, 8), ("f32", torch.float32, 4)):
        p = "d" if dt == "f64" else "s"
        for D in map(int, args.gemm_dims.split(",")):
            for name, m, n, k in gemm_types(D):
                nbytes = (m * k + k * n + m * n) * es
                copies = max(1, min(8, -(-2 * L2_BYTES // nbytes))) if nbytes < 2 * L2_BYTES else 1
                sets = [(fill(m * k, 7.0, tdt), fill(k * n, 3.0, tdt), torch.zeros(m * n, dtype=tdt, device="cuda"))
                        for _ in range(copies)]
                calls = [(lambda a=a, b=b, c=c: ctx.gemm(dt, m, n, k, 1.0, a, m, b, k, 0.0, c, m)) for a, b, c in sets]
                ms = time_calls(ctx, calls, args.iters)
                kern = ctx.last_kernel()
                flops = 2 * m * n * k + m * n
                gf = flops / ms * 1e-6
                pk = pipe["f64"] if dt == "f64" else (pipe["f32_tc"] if "tcgen05" in kern or "tf32" in kern else pipe["f32_ffma"])
                bound = min(pk, flops / nbytes * hbm)
                rows.append((f"{p}gemm", name, f"{m}x{n}x{k}", ms * 1e3, gf, nbytes / ms * 1e-6,
                             "pipe" if bound == pk else "hbm", bound, gf / bound, kern, copies))
                del sets, calls
                torch.cuda.empty_cache()
        for D in map(int, args.gemv_dims.split(",")):
            for name, m, n in gemv_types(D):
                nbytes = (m * n + n + m) * es
                copies = max(1, min(8, -(-2 * L2_BYTES // nbytes))) if nbytes < 2 * L2_BYTES else 1
                sets = [(fill(m * n, 7.0, tdt), fill(n, 3.0, tdt), torch.zeros(m, dtype=tdt, device="cuda"))
                        for _ in range(copies)]
                calls = [(lambda a=a, x=x, y=y: ctx.gemv(dt, m, n, 1.0, a, m, x, 1, 0.0, y, 1)) for a, x, y in sets]
                ms = time_calls(ctx, calls, args.iters)
                flops = 2 * m * n + m
                gf = flops / ms * 1e-6
                bound = flops / nbytes * hbm
                rows.append((f"{p}gemv", name, f"{m}x{n}", ms * 1e3, gf, nbytes / ms * 1e-6, "hbm", bound,
                             gf / bound, ctx.last_kernel(), copies))
                del sets, calls
                torch.cuda.empty_cache()
    hdr = ("| op | problem type | shape | us/call | GFLOP/s | GB/s (algorithmic) | bound | roofline GFLOP/s | frac | kernel | L2-rotation copies |\n"
           "|---|---|---|---|---|---|---|---|---|---|---|\n")
    body = "".join(f"| {r[0]} | {r[1]} | {r[2]} | {r[3]:.1f} | {r[4]:.0f} | {r[5]:.0f} | {r[6]} | {r[7]:.0f} | {r[8]:.2f} | `{r[9]}` | {r[10]} |\n"
                   for r in rows)
    head = (f"Denominators: HBM {hbm:.0f} GB/s (MEASURED_PEAKS.json), FP64 pipe {pipe['f64']:.0f} GFLOP/s (probe), "
            f"FFMA {pipe['f32_ffma']:.0f} GFLOP/s (probe), 3xTF32 tensor {pipe['f32_tc']:.0f} GFLOP/s (bf16/2/3).\n\n")
    out = head + hdr + body
    print(out)
    if args.md:
        open(args.md, "w").write(out)


if __name__ == "__main__":
    main()
