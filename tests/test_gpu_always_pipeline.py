"""Always mode above the one-shot limit: the dependency-driven block pipeline of
b200_?gemm_offload (K-chunks of A x column slabs of B / C, merged K runs, slab downloads) and
the column-slab streaming of b200_?gemv_offload, against the Once-mode result of the same
inputs (itself pinned to the oracle elsewhere) and against the oracle on sampled blocks."""
import numpy as np
import pytest

from oracle_lib import max_rel_err
from test_gpu_parity import run_gemm, run_gemv, TOL, NP

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt,shape", [("f64", (2048, 2048, 2048)), ("f64", (1536, 2300, 4100)), ("f64", (4096, 4096, 3000)),
                                      ("f32", (2048, 2048, 2048)), ("f32", (3000, 2500, 1100)), ("f64", (1030, 4100, 1026)),
                                      # compute-bound regime (FP64 from ~5500^3): interleaved A / B uploads, 6 x 6 pieces,
                                      # K-chunk runs joined into one GEMM -- the path bench.py's e2e number takes
                                      ("f64", (6144, 6144, 6144)), ("f64", (5900, 6100, 5700))],
                         ids=lambda v: v if isinstance(v, str) else "x".join(map(str, v)))
def test_gemm_block_pipeline(blob, ctx, oracle, dt, shape):
    m, n, k = shape
    rng = np.random.default_rng(m + n + k)
    A = (rng.integers(0, 100, m * k) / 7.0).astype(NP[dt])
    B = (rng.integers(0, 100, k * n) / 3.0).astype(NP[dt])
    once, _ = run_gemm(blob, ctx, dt, "once", m, n, k, A, B)
    always, kern = run_gemm(blob, ctx, dt, "always", m, n, k, A, B, iters=2)
    assert max_rel_err(always, once) <= TOL[dt], (dt, shape, kern)
    Am, Bm, Cm = A.reshape(k, m), B.reshape(n, k), always.reshape(n, m)
    bs = 24
    for (i0, j0) in ((0, 0), (m - bs, n - bs), (m // 2, n // 2 + 5), (7, n - bs)):
        a_cm = np.ascontiguousarray(Am[:, i0:i0 + bs]).reshape(-1)
        b_cm = np.ascontiguousarray(Bm[j0:j0 + bs, :]).reshape(-1)
        want = np.zeros(bs * bs, dtype=NP[dt])
        oracle.gemm(dt, bs, bs, k, 1.0, a_cm, bs, b_cm, k, 0.0, want, bs)
        got = np.ascontiguousarray(Cm[j0:j0 + bs, i0:i0 + bs]).reshape(-1)
        assert max_rel_err(got, want) <= TOL[dt], (dt, shape, kern, (i0, j0))


def test_gemm_block_pipeline_beta(blob, ctx, oracle):
    """beta != 0: C is uploaded first and only the first K-chunk of a slab applies beta."""
    m, n, k = 2048, 1536, 2200
    rng = np.random.default_rng(77)
    A = rng.random(m * k)
    B = rng.random(k * n)
    C0 = rng.random(m * n) * 100
    once, _ = run_gemm(blob, ctx, "f64", "once", m, n, k, A, B, C0=C0, alpha=0.5, beta=-1.5)
    always, kern = run_gemm(blob, ctx, "f64", "always", m, n, k, A, B, C0=C0, alpha=0.5, beta=-1.5)
    assert max_rel_err(always, once) <= 1e-11, kern  # some cancellation against -1.5 C0


@pytest.mark.parametrize("dt,shape", [("f64", (4096, 4096)), ("f32", (8192, 3000)), ("f64", (515, 30000))],
                         ids=lambda v: v if isinstance(v, str) else "x".join(map(str, v)))
def test_gemv_slab_streaming(blob, ctx, oracle, dt, shape):
    m, n = shape
    rng = np.random.default_rng(m + n)
    A = (rng.integers(0, 100, m * n) / 7.0).astype(NP[dt])
    x = (rng.integers(0, 100, n) / 3.0).astype(NP[dt])
    want = np.zeros(m, dtype=NP[dt])
    oracle.gemv(dt, m, n, 1.0, A, m, x, 1, 0.0, want, 1)
    got, kern = run_gemv(blob, ctx, dt, "always", m, n, A, x, iters=2)
    assert max_rel_err(got, want) <= TOL[dt], (dt, shape, kern)
