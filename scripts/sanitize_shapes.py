#!/usr/bin/env python
"""A handful of ragged shapes through every tensor-core schedule, meant to be run under
`compute-sanitizer --tool memcheck` (results are also checked against torch fp64 matmul)."""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
ctx = blob.Context(0)
cases = [("f32", 516, 772, 264, {"sgemm": "tc", "sgemm_pair": "on"}),      # pair kernel, ragged m/n/k
         ("f32", 260, 132, 1100, {"sgemm": "tc", "sgemm_pair": "off"}),    # single CTA + split-K
         ("f32", 1028, 260, 72, {"sgemm": "tc", "sgemm_pair": "off"}),     # single CTA, no split
         ("f64", 1700, 1650, 300, {}),                                      # TMA persistent + stream-K
         ("f64", 516, 520, 4100, {}),                                       # TMA split-K (tall K)
         ("f64", 300, 260, 200, {}),                                        # 32x32 DMMA split-K
         ("f64", 32, 32, 4100, {})]                                         # M = N = 32, long K
for dt, m, n, k, opts in cases:
    for key, val in opts.items():
        ctx.set_option(key, val)
    tdt = torch.float64 if dt == "f64" else torch.float32
    lda, ldb, ldc = m + (4 - m % 4) % 4, k + (4 - k % 4) % 4, m + 1
    A = torch.rand(lda * k, dtype=tdt, device="cuda")
    B = torch.rand(ldb * n, dtype=tdt, device="cuda")
    C = torch.rand(ldc * n, dtype=tdt, device="cuda")
    C0 = C.clone()
    torch.cuda.synchronize()
    ctx.gemm(dt, m, n, k, 1.25, A, lda, B, ldb, 0.5, C, ldc)
    ctx.sync()
    ref = 1.25 * (A.view(k, lda)[:, :m].double().T @ B.view(n, ldb)[:, :k].double().T) + 0.5 * C0.view(n, ldc)[:, :m].double().T
    got = C.view(n, ldc)[:, :m].double().T
    err = ((got - ref).abs() / ref.abs()).max().item()
    pad_ok = torch.equal(C.view(n, ldc)[:, m:], C0.view(n, ldc)[:, m:])
    print(f"{dt} {m}x{n}x{k} kernel={ctx.last_kernel()} max_rel_err={err:.2e} padding_untouched={pad_ok}")
    assert err < (1e-5 if dt == "f32" else 1e-12) and pad_ok
    for key in opts:
        ctx.set_option(key, "auto")
ctx.close()
print("ok")
