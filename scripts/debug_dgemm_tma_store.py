#!/usr/bin/env python
"""Debug: DGEMM beta=0 -> beta=1 chains against torch FP64, per library build / store option.
usage: LIB=<path to .so> STORE=on|off debug_dgemm_tma_store.py"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
if os.environ.get("LIB"):
    blob.LIB_PATH = os.environ["LIB"]
store = os.environ.get("STORE", "on")
ctx = blob.Context(0)
try:
    ctx.set_option("dgemm_tma_store", store)
except Exception as ex:
    print("no dgemm_tma_store option in this build:", str(ex)[:60])
def fill(n, div, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 100, (n,), generator=g, device="cuda").to(torch.float64) / div
def run(m, n, k, chain, A, B, sync_between=False):
    C = torch.full((m * n,), -1.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    kc = k // chain
    for rep in range(2):
        for c in range(chain):
            ctx.gemm("f64", m, n, kc, 1.0, A[c * kc * m:], m, B[c * kc:], k, 0.0 if c == 0 else 1.0, C, m)
            if sync_between:
                ctx.sync()
    ctx.sync()
    return C
for (m, n, k, chain) in [(4096, 4096, 3072, 1), (4096, 4096, 3072, 3), (6144, 1024, 6144, 6), (2048, 2048, 2048, 4), (4096, 4096, 1024, 2)]:
    A, B = fill(m * k, 7.0, 1), fill(k * n, 3.0, 2)
    ref = (B.view(n, k) @ A.view(k, m)).reshape(-1)
    for sb in (False, True):
        Cs = [run(m, n, k, chain, A, B, sb) for _ in range(2)]
        errs = [((c - ref).abs() / ref.abs()).max().item() for c in Cs]
        nb = int((Cs[0] != Cs[1]).sum().item())
        print(f"lib={os.path.basename(blob.LIB_PATH)} store={store} {m}x{n}x{k} chain={chain} sync_between={sb} kernel={ctx.last_kernel()} "
              f"err_vs_torch={errs[0]:.2e},{errs[1]:.2e} run-to-run mismatches={nb}", flush=True)
