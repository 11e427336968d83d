"""Stream-K schedule of the CTA-pair tcgen05 SGEMM kernel (partial last wave of 256 x 256 tiles
cut into equal runs of K = 128 accumulation chunks; partial tiles combined in contributor order
by the last CTA to arrive): parity with the oracle on sampled blocks at the FP32 bar, agreement
with the whole-tile schedule, determinism, alpha / beta, ragged edges."""
import numpy as np
import pytest

from oracle_lib import max_rel_err

pytestmark = pytest.mark.gpu


def _fill(torch, count, seed, div):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.randint(0, 100, (count,), generator=g, device="cuda").to(torch.float32) / div


@pytest.mark.parametrize("shape,alpha,beta", [((2304, 2304, 2304), 1.0, 0.0), ((2500, 2700, 3300), 1.0, 0.0),
                                              ((4096, 4096, 4096), 1.0, 0.0), ((2304, 2308, 2052), 1.5, -0.5)],
                         ids=lambda v: "x".join(map(str, v)) if isinstance(v, tuple) else str(v))
def test_sgemm_pair_streamk(blob, ctx, oracle, shape, alpha, beta):
    import torch
    m, n, k = shape
    lda, ldb, ldc = (m + 3) // 4 * 4, (k + 3) // 4 * 4, m + 1
    A = _fill(torch, lda * k, 31, 7.0)
    B = _fill(torch, ldb * n, 32, 3.0)
    C0 = _fill(torch, ldc * n, 33, 0.001)
    C, C2, C3 = C0.clone(), C0.clone(), C0.clone()
    torch.cuda.synchronize()  # inputs are produced on torch's stream, the library runs on its own
    ctx.gemm("f32", m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
    ctx.sync()
    kern = ctx.last_kernel()
    assert kern.endswith("pair_256x256x32_streamk"), kern
    ctx.gemm("f32", m, n, k, alpha, A, lda, B, ldb, beta, C2, ldc)
    ctx.set_option("sgemm_streamk", "off")
    try:
        ctx.gemm("f32", m, n, k, alpha, A, lda, B, ldb, beta, C3, ldc)
        ctx.sync()
        assert "pair_256x" in ctx.last_kernel() and not ctx.last_kernel().endswith("streamk"), ctx.last_kernel()
    finally:
        ctx.set_option("sgemm_streamk", "on")
    assert torch.equal(C, C2)  # contributor order, not arrival order
    Cv, C3v, C0v = C.view(n, ldc), C3.view(n, ldc), C0.view(n, ldc)
    assert torch.equal(Cv[:, m:], C0v[:, m:])  # padding rows untouched
    rel = ((Cv[:, :m] - C3v[:, :m]).abs() / C3v[:, :m].abs().clamp_min(1e-30)).max().item()
    assert rel <= 1e-5, rel
    Ah, Bh, Ch, C0h = A.view(k, lda).cpu().numpy(), B.view(n, ldb).cpu().numpy(), Cv.cpu().numpy(), C0v.cpu().numpy()
    bs = 40
    for (i0, j0) in ((0, 0), (m - bs, n - bs), (m - bs, 0), (m // 2 + 3, n // 2 - 9), (250, n - bs), (m - bs, 1000)):
        a_cm = np.ascontiguousarray(Ah[:, i0:i0 + bs]).reshape(-1)
        b_cm = np.ascontiguousarray(Bh[j0:j0 + bs, :k]).reshape(-1)
        want = np.ascontiguousarray(C0h[j0:j0 + bs, i0:i0 + bs]).reshape(-1).copy()
        oracle.gemm("f32", bs, bs, k, alpha, a_cm, bs, b_cm, k, beta, want, bs)
        got = np.ascontiguousarray(Ch[j0:j0 + bs, i0:i0 + bs]).reshape(-1)
        assert max_rel_err(got, want) <= 1e-5, (shape, kern, (i0, j0))
