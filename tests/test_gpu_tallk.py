"""The cluster / distributed-shared-memory split-K kernel for one small tile of C and a long K
(csrc/gemm_tallk.cu): the reference's M = N = 32, K = d problem type (include/doGemm.hh:171-201).
Parity with the oracle at BASELINE.json's tolerances, ragged and unaligned operands, beta != 0,
deterministic summation order, and proof that this kernel is the one that ran."""
import numpy as np
import pytest

from oracle_lib import max_rel_err
from test_gpu_parity import run_gemm

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-12, "f32": 1e-5}
NP = {"f64": np.float64, "f32": np.float32}

SHAPES = [(32, 32, 1024), (32, 32, 2048), (32, 32, 8192), (32, 32, 32768), (32, 32, 1027), (17, 29, 5000),
          (1, 1, 4096), (32, 1, 9000), (5, 32, 1500), (32, 32, 70000)]


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_tallk_matches_the_harness_problem(blob, ctx, oracle, shape, dt):
    m, n, k = shape
    A, B, _ = oracle.init_gemm(dt, m, n, k)
    want = oracle.harness_gemm(dt, m, n, k)
    got, kern = run_gemm(blob, ctx, dt, "once", m, n, k, A, B, iters=3)  # repeated calls: self-resetting counters
    assert kern.endswith("tallk_32x32_cluster8"), kern
    assert max_rel_err(got, want) <= TOL[dt], (shape, dt, kern)
    got2, _ = run_gemm(blob, ctx, dt, "once", m, n, k, A, B, iters=1)
    assert np.array_equal(got, got2)  # summation order fixed by indices, not by arrival


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_tallk_alpha_beta_and_leading_dimensions(blob, ctx, oracle, dt):
    rng = np.random.default_rng(21)
    m, n, k = 30, 27, 3001
    lda, ldb, ldc = 33, 3005, 37  # odd strides: no 16-byte alignment anywhere
    A = rng.random(lda * k).astype(NP[dt])
    B = rng.random(ldb * n).astype(NP[dt])
    C0 = (rng.random(ldc * n) * 100).astype(NP[dt])
    want = C0.copy()
    oracle.gemm(dt, m, n, k, 1.25, A, lda, B, ldb, -0.5, want, ldc)
    got, kern = run_gemm(blob, ctx, dt, "once", m, n, k, A, B, C0=C0, alpha=1.25, beta=-0.5, lda=lda, ldb=ldb, ldc=ldc)
    assert kern.endswith("tallk_32x32_cluster8"), kern
    g, w, c0 = got.reshape(n, ldc), want.reshape(n, ldc), C0.reshape(n, ldc)
    assert np.array_equal(g[:, m:], c0[:, m:])  # padding rows untouched
    assert max_rel_err(g[:, :m].ravel(), w[:, :m].ravel()) <= TOL[dt] * 10


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_tallk_agrees_with_the_generic_split_k_path(blob, ctx, oracle, dt):
    """Same problem through the kernel it replaced (DMMA / FFMA split-K): both within tolerance
    of the oracle and of each other."""
    m, n, k = 32, 32, 8192
    A, B, _ = oracle.init_gemm(dt, m, n, k)
    got, kern = run_gemm(blob, ctx, dt, "once", m, n, k, A, B)
    ctx.set_option("tallk", "off")
    try:
        old, kern_old = run_gemm(blob, ctx, dt, "once", m, n, k, A, B)
    finally:
        ctx.set_option("tallk", "on")
    assert "tallk" in kern and "tallk" not in kern_old, (kern, kern_old)
    assert max_rel_err(got, old) <= 2 * TOL[dt]


def test_tallk_always_mode_and_below_threshold(blob, ctx, oracle):
    """Always mode reaches the same kernel through the transfer engine; K < 1024 keeps the old path."""
    A, B, _ = oracle.init_gemm("f64", 32, 32, 4096)
    want = oracle.harness_gemm("f64", 32, 32, 4096)
    got, kern = run_gemm(blob, ctx, "f64", "always", 32, 32, 4096, A, B, iters=2)
    assert "tallk" in kern, kern
    assert max_rel_err(got, want) <= 1e-12
    A, B, _ = oracle.init_gemm("f64", 32, 32, 512)
    _, kern = run_gemm(blob, ctx, "f64", "once", 32, 32, 512, A, B)
    assert "tallk" not in kern, kern
