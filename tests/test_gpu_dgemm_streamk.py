"""Stream-K schedule of the TMA-fed DGEMM kernel (partial last wave cut into equal k-tile
segments, partial tiles added in contributor order by the last arriver): parity with the
oracle on sampled C blocks at the FP64 bar, agreement with the whole-tile schedule, and
run-to-run determinism."""
import numpy as np
import pytest

from oracle_lib import max_rel_err

pytestmark = pytest.mark.gpu


def _fill(torch, count, seed, div):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 100, (count,), generator=g, device="cuda").to(torch.float64) / div


@pytest.mark.parametrize("shape", [(2048, 2048, 2048), (1700, 1650, 1100), (8192, 512, 512), (1920, 2176, 4100)],
                         ids=lambda s: "x".join(map(str, s)))
def test_dgemm_streamk(blob, ctx, oracle, shape):
    import torch
    m, n, k = shape
    lda, ldb, ldc = m + (m & 1), k + (k & 1), m
    A = _fill(torch, lda * k, 11, 7.0)
    B = _fill(torch, ldb * n, 12, 3.0)
    C = torch.zeros(ldc * n, dtype=torch.float64, device="cuda")
    C2 = torch.zeros_like(C)
    C3 = torch.zeros_like(C)
    torch.cuda.synchronize()  # inputs are produced on torch's stream, the library runs on its own
    ctx.gemm("f64", m, n, k, 1.0, A, lda, B, ldb, 0.0, C, ldc)
    ctx.sync()
    kern = ctx.last_kernel()
    assert kern.endswith("streamk"), kern
    ctx.gemm("f64", m, n, k, 1.0, A, lda, B, ldb, 0.0, C2, ldc)
    ctx.set_option("dgemm_streamk", "off")
    try:
        ctx.gemm("f64", m, n, k, 1.0, A, lda, B, ldb, 0.0, C3, ldc)
        ctx.sync()
        assert not ctx.last_kernel().endswith("streamk")
    finally:
        ctx.set_option("dgemm_streamk", "on")
    assert torch.equal(C, C2)  # deterministic: contributor order, not arrival order
    rel = ((C - C3).abs() / C3.abs().clamp_min(1e-300)).max().item()
    assert rel <= 1e-12, rel
    # sampled blocks against the oracle (corners, a block straddling tiles, the last ragged tile)
    Ah = A.view(k, lda).cpu().numpy()
    Bh = B.view(n, ldb).cpu().numpy()
    Ch = C.view(n, ldc).cpu().numpy()
    bs = 40
    for (i0, j0) in ((0, 0), (m - bs, n - bs), (0, n - bs), (m // 2 - 7, n // 2 + 3), (m - bs, 0), (120, 250)):
        a_cm = np.ascontiguousarray(Ah[:, i0:i0 + bs]).reshape(-1)          # (i, p) at p*bs + i
        b_cm = np.ascontiguousarray(Bh[j0:j0 + bs, :k]).reshape(-1)         # (p, j) at j*k + p
        want = np.zeros(bs * bs, dtype=np.float64)
        oracle.gemm("f64", bs, bs, k, 1.0, a_cm, bs, b_cm, k, 0.0, want, bs)
        got = np.ascontiguousarray(Ch[j0:j0 + bs, i0:i0 + bs]).reshape(-1)
        assert max_rel_err(got, want) <= 1e-12, (shape, kern, (i0, j0))


@pytest.mark.parametrize("shape,expect", [((1700, 1650, 1100), "streamk"), ((512, 512, 8192), "splitk"),
                                          ((2048, 2048, 96), "128x128x32")],
                         ids=["streamk", "splitk", "persistent"])
def test_dgemm_tma_alpha_beta(blob, ctx, oracle, shape, expect):
    """alpha / beta through every schedule of the TMA kernel: C = 1.5 A B - 0.25 C0, checked on
    sampled blocks against the oracle (the accumulate blocks of the Always-mode pipeline use
    beta = 1 on exactly these paths)."""
    import torch
    m, n, k = shape
    A = _fill(torch, m * k, 21, 7.0)
    B = _fill(torch, k * n, 22, 3.0)
    C0 = _fill(torch, m * n, 23, 11.0)
    C = C0.clone()
    torch.cuda.synchronize()
    ctx.gemm("f64", m, n, k, 1.5, A, m, B, k, -0.25, C, m)
    ctx.sync()
    kern = ctx.last_kernel()
    assert kern.endswith(expect), kern
    Ah, Bh, Ch, C0h = A.view(k, m).cpu().numpy(), B.view(n, k).cpu().numpy(), C.view(n, m).cpu().numpy(), \
        C0.view(n, m).cpu().numpy()
    bs = 36
    for (i0, j0) in ((0, 0), (m - bs, n - bs), (m // 2, n // 3), (130, n - bs)):
        a_cm = np.ascontiguousarray(Ah[:, i0:i0 + bs]).reshape(-1)
        b_cm = np.ascontiguousarray(Bh[j0:j0 + bs, :]).reshape(-1)
        want = np.ascontiguousarray(C0h[j0:j0 + bs, i0:i0 + bs]).reshape(-1).copy()
        oracle.gemm("f64", bs, bs, k, 1.5, a_cm, bs, b_cm, k, -0.25, want, bs)
        got = np.ascontiguousarray(Ch[j0:j0 + bs, i0:i0 + bs]).reshape(-1)
        assert max_rel_err(got, want) <= 1e-12, (shape, kern, (i0, j0))


@pytest.mark.parametrize("pad", ["off", "on"])
@pytest.mark.parametrize("shape", [(4096, 4096, 1024), (6144, 1024, 1024), (4096, 4096, 96)], ids=lambda s: "x".join(map(str, s)))
def test_dgemm_accumulate_blocks_are_exact(blob, ctx, shape, pad):
    """C += A * B (beta = 1: the accumulate blocks of the Always-mode pipeline and of the multi-GPU
    K-chunk loop) must equal C_old + (A * B by the beta = 0 call) to the last bit -- the update
    is one FMA per element -- on every repetition.  Regression for a stage-release race in the TMA
    pipeline: the mbarrier arrive that frees a stage could issue while the last fragment loads of
    that stage were still in flight; with the LSU queue backed up by the beta != 0 epilogue a
    refill overtook them and single DMMA steps ran on the wrong k-tile.  "dgemm_smem_pad" = on
    launches with the maximum shared-memory carve-out (minimum L1), where the unfixed kernel
    failed within a few launches."""
    import torch
    m, n, k = shape
    A = _fill(torch, m * k, 51, 7.0)
    B = _fill(torch, k * n, 52, 3.0)
    Cold = _fill(torch, m * n, 53, 0.001)
    P = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    ctx.set_option("dgemm_smem_pad", pad)
    try:
        ctx.gemm("f64", m, n, k, 1.0, A, m, B, k, 0.0, P, m)
        ctx.sync()
        assert ctx.last_kernel().startswith("dgemm_dmma_tma"), ctx.last_kernel()
        want = Cold + P
        for rep in range(6):
            C = Cold.clone()
            torch.cuda.synchronize()
            ctx.gemm("f64", m, n, k, 1.0, A, m, B, k, 1.0, C, m)
            ctx.sync()
            bad = int((C != want).sum().item())
            assert bad == 0, (shape, pad, rep, bad)
    finally:
        ctx.set_option("dgemm_smem_pad", "off")
