"""Ordering and BLAS-argument semantics of the C ABI that the harness never stresses but a
generic caller of include/b200blas.h relies on (round-1 advisor findings):

* a download is ordered before later kernels that overwrite the buffer (the reference gets this
  from the legacy default stream, cuBLAS/gemm.hh:60-62,149-150);
* the Always-mode engine honours leading dimensions larger than the row count: host padding
  is never overwritten and a minimally sized caller buffer is never over-read;
* B200_UPLOAD_C / option "upload_c" (the reference's dead C upload, cuBLAS/gemm.hh:130-131)
  changes traffic, not results.
"""
import numpy as np
import pytest

from oracle_lib import max_rel_err
from test_gpu_parity import Operand, NP, TOL

pytestmark = pytest.mark.gpu


def test_download_is_ordered_before_a_later_kernel(blob, ctx, oracle):
    """d2h(C) then, with NO sync in between, a second GEMM into the same C: the host must
    receive the first result."""
    m = n = 3072
    k1, k2 = 64, 32
    rng = np.random.default_rng(1)
    A1, B1 = rng.random(m * k1), rng.random(k1 * n)
    A2, B2 = rng.random(m * k2) + 5.0, rng.random(k2 * n) + 5.0
    a1, b1 = Operand(blob, ctx, 1, "f64", m * k1), Operand(blob, ctx, 1, "f64", k1 * n)
    a2, b2 = Operand(blob, ctx, 1, "f64", m * k2), Operand(blob, ctx, 1, "f64", k2 * n)
    c = Operand(blob, ctx, 1, "f64", m * n)
    try:
        for op, v in ((a1, A1), (b1, B1), (a2, A2), (b2, B2)):
            op.np[:] = v
            ctx.h2d(op.dev, op.host, op.nbytes, 0)
        want = np.zeros(m * n)
        oracle.gemm("f64", m, n, k1, 1.0, A1, m, B1, k1, 0.0, want, m)
        for _ in range(3):
            c.np[:] = -1.0
            ctx.gemm("f64", m, n, k1, 1.0, a1.dev, m, b1.dev, k1, 0.0, c.dev, m)
            ctx.d2h(c.host, c.dev, c.nbytes, 2)                                   # 72 MB, ~1.4 ms on the wire
            ctx.gemm("f64", m, n, k2, 1.0, a2.dev, m, b2.dev, k2, 0.0, c.dev, m)  # overwrites C in ~0.2 ms
            ctx.sync()
            assert max_rel_err(c.np, want) <= 1e-12
    finally:
        for op in (a1, b1, a2, b2, c):
            op.free()


@pytest.mark.parametrize("shape", [(300, 200, 100), (1500, 1100, 900), (2500, 2300, 2100)],
                         ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_always_mode_respects_leading_dimensions(blob, ctx, oracle, dt, shape):
    """ld > rows through b200_?gemm_offload (zero-copy / one-shot / block-pipeline sizes): the
    result is right, the host padding of C keeps its sentinel, and A / B buffers hold exactly
    ld * (cols - 1) + rows elements (an over-read would leave the pinned block)."""
    m, n, k = shape
    lda, ldb, ldc = m + 3, k + 5, m + 7
    rng = np.random.default_rng(sum(shape))
    t = NP[dt]
    A = rng.random(lda * (k - 1) + m).astype(t)
    B = rng.random(ldb * (n - 1) + k).astype(t)
    a, b = Operand(blob, ctx, 0, dt, A.size), Operand(blob, ctx, 0, dt, B.size)
    c = Operand(blob, ctx, 0, dt, ldc * n)
    try:
        a.np[:], b.np[:] = A, B
        c.np[:] = -7.0
        ctx.gemm_offload(dt, m, n, k, 1.0, a.host, a.dev, lda, b.host, b.dev, ldb, 0.0, c.host, c.dev, ldc)
        got = c.np.reshape(n, ldc)
        want = np.zeros(ldc * n, dtype=t)
        oracle.gemm(dt, m, n, k, 1.0, A, lda, B, ldb, 0.0, want, ldc)
        assert max_rel_err(got[:, :m], want.reshape(n, ldc)[:, :m]) <= TOL[dt]
        assert (got[:, m:] == -7.0).all(), "host padding rows of C were overwritten"
    finally:
        a.free(); b.free(); c.free()


@pytest.mark.parametrize("shape", [(200, 150, 100), (2048, 2048, 2048)], ids=lambda s: "x".join(map(str, s)))
def test_upload_c_option_changes_traffic_not_results(blob, oracle, shape):
    m, n, k = shape
    A, B, _ = oracle.init_gemm("f64", m, n, k)
    outs = []
    for opt in ("off", "on"):
        with blob.Context() as c2:
            c2.set_option("upload_c", opt)
            a, b, c = (Operand(blob, c2, 0, "f64", x) for x in (m * k, k * n, m * n))
            a.np[:], b.np[:] = A, B
            c.np[:] = np.nan     # beta == 0: whatever is uploaded must not leak into the result
            c2.gemm_offload("f64", m, n, k, 1.0, a.host, a.dev, m, b.host, b.dev, k, 0.0, c.host, c.dev, m)
            outs.append(c.np.copy())
            a.free(); b.free(); c.free()
    assert not np.isnan(outs[1]).any()
    assert np.array_equal(outs[0], outs[1])
