"""The tcgen05 / TMEM 3xTF32 SGEMM path: parity with the oracle at the FP32 bar
(<= 1e-5 element-wise), including ragged tiles (TMA zero fill), long K (FP32
accumulation in TMEM) and beta != 0, and proof that the tensor-core kernel is
the one that ran."""
import numpy as np
import pytest

from oracle_lib import max_rel_err
from test_gpu_parity import run_gemm

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["fused", "presplit"])
def force_tensor_core_path(ctx, request):
    """Small shapes default to the FFMA path (launch-latency crossover ~512^3);
    these tests pin the tcgen05 kernel itself -- once with the operand split fused into the
    GEMM (converter warps, the default) and once with the separate pre-split pass."""
    ctx.set_option("sgemm", "tc")
    ctx.set_option("sgemm_fused", "on" if request.param == "fused" else "off")
    yield request.param
    ctx.set_option("sgemm_fused", "auto")
    ctx.set_option("sgemm", "auto")

SHAPES = [
    (128, 128, 256), (256, 256, 256), (384, 512, 96), (640, 384, 1000), (1024, 1024, 64),
    (260, 132, 200), (1028, 516, 260), (2048, 2048, 32), (512, 64, 4096), (64, 512, 4100),
    (128, 128, 8192), (256, 256, 32768), (4096, 64, 4096), (1024, 1024, 1024), (2052, 1028, 516),
]


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_sgemm_tensor_core_path(blob, ctx, oracle, shape, force_tensor_core_path):
    m, n, k = shape
    A, B, _ = oracle.init_gemm("f32", m, n, k)
    want = oracle.harness_gemm("f32", m, n, k)
    got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, iters=2)
    assert "tcgen05" in kern, kern
    assert ("fused" in kern) == (force_tensor_core_path == "fused"), kern
    err = max_rel_err(got, want)
    assert err <= 1e-5, (shape, kern, err)


def test_sgemm_tensor_core_alpha_beta(blob, ctx, oracle):
    rng = np.random.default_rng(3)
    m, n, k = 300, 260, 128
    lda, ldb, ldc = 304, 128, 301
    A = rng.random(lda * k).astype(np.float32)
    B = rng.random(ldb * n).astype(np.float32)
    C0 = rng.random(ldc * n).astype(np.float32)
    want = C0.copy()
    oracle.gemm("f32", m, n, k, 1.5, A, lda, B, ldb, 0.5, want, ldc)
    got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, C0=C0, alpha=1.5, beta=0.5,
                         lda=lda, ldb=ldb, ldc=ldc)
    assert "tcgen05" in kern, kern
    assert max_rel_err(got, want) <= 1e-5


def test_sgemm_signed_inputs_absolute_error(blob, ctx, oracle):
    """With cancellation the relative bar is meaningless; the 3xTF32 absolute error
    must stay at FP32 level relative to |A||B|."""
    rng = np.random.default_rng(9)
    m = n = k = 512
    A = (rng.random(m * k) - 0.5).astype(np.float32)
    B = (rng.random(k * n) - 0.5).astype(np.float32)
    want = np.zeros(m * n, np.float32)
    oracle.gemm("f32", m, n, k, 1.0, A, m, B, k, 0.0, want, m)
    bound = np.zeros(m * n, np.float32)
    oracle.gemm("f32", m, n, k, 1.0, np.abs(A), m, np.abs(B), k, 0.0, bound, m)
    got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B)
    assert "tcgen05" in kern, kern
    assert float(np.max(np.abs(got.astype(np.float64) - want) / bound)) <= 1e-5


# ---- CTA-pair (cta_group::2) 256 x 256 kernel, forced ------------------------------
PAIR_SHAPES = [
    (256, 256, 32), (256, 256, 256), (512, 512, 96), (768, 1280, 1000), (260, 300, 200),
    (1028, 516, 260), (2048, 2048, 32), (512, 256, 4100), (256, 256, 32768), (4096, 256, 4096),
    (1024, 1024, 1024), (2052, 1028, 516), (300, 2304, 72), (5000, 3000, 136),
]


@pytest.fixture
def force_pair(ctx):
    ctx.set_option("sgemm_pair", "on")
    yield
    ctx.set_option("sgemm_pair", "auto")


@pytest.mark.parametrize("shape", PAIR_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_sgemm_pair_kernel(blob, ctx, oracle, force_pair, shape):
    m, n, k = shape
    A, B, _ = oracle.init_gemm("f32", m, n, k)
    want = oracle.harness_gemm("f32", m, n, k)
    got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, iters=2)
    assert "pair_256x256" in kern, kern
    err = max_rel_err(got, want)
    assert err <= 1e-5, (shape, kern, err)


def test_sgemm_pair_alpha_beta_ld(blob, ctx, oracle, force_pair):
    rng = np.random.default_rng(4)
    m, n, k = 700, 520, 200
    lda, ldb, ldc = 704, 200, 701
    A = rng.random(lda * k).astype(np.float32)
    B = rng.random(ldb * n).astype(np.float32)
    C0 = rng.random(ldc * n).astype(np.float32)
    want = C0.copy()
    oracle.gemm("f32", m, n, k, 1.5, A, lda, B, ldb, 0.5, want, ldc)
    got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, C0=C0, alpha=1.5, beta=0.5,
                         lda=lda, ldb=ldb, ldc=ldc)
    assert "pair_256x256" in kern, kern
    assert max_rel_err(got, want) <= 1e-5


def test_sgemm_pair_matches_single_cta_bitwise(blob, ctx, oracle, force_pair):
    """Same products, same chunked accumulation order: the two tile configs must agree
    to the last bit (a layout or barrier bug in the pair kernel would not)."""
    m, n, k = 2048, 1536, 264  # 192 single-CTA tiles: more than one wave, so no split-K there
    A, B, _ = oracle.init_gemm("f32", m, n, k)
    got2, kern2 = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B)
    ctx.set_option("sgemm_pair", "off")
    got1, kern1 = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B)
    assert "pair" in kern2 and kern1.endswith("128x128x32"), (kern1, kern2)
    assert np.array_equal(got1, got2)


def test_sgemm_fused_split_leaves_no_launch_and_no_workspace_copy(blob, ctx, oracle, force_tensor_core_path):
    """Fused: ONE launch per SGEMM (no split_tf32_kernel); pre-split: two."""
    m, n, k = 1024, 768, 512
    A, B, _ = oracle.init_gemm("f32", m, n, k)
    l0 = ctx.launch_count()
    _, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, iters=1)
    per_call = ctx.launch_count() - l0
    assert per_call == (1 if force_tensor_core_path == "fused" else 2), (kern, per_call)


def test_sgemm_fused_handles_denormals_zeros_and_negative_values(blob, ctx, oracle, force_tensor_core_path):
    """lo = RN_tf32(x - trunc_tf32(x)) must be exact-enough for every sign / magnitude mix."""
    rng = np.random.default_rng(12)
    m, n, k = 256, 256, 200
    A = (rng.standard_normal(m * k) * 10.0 ** rng.integers(-6, 6, m * k)).astype(np.float32)
    B = (rng.standard_normal(k * n) * 10.0 ** rng.integers(-6, 6, k * n)).astype(np.float32)
    A[::7] = 0.0
    B[::5] = 0.0
    A[3::11] = np.float32(1e-41)  # denormal
    want = np.zeros(m * n, np.float32)
    oracle.gemm("f32", m, n, k, 1.0, A, m, B, k, 0.0, want, m)
    bound = np.zeros(m * n, np.float32)
    oracle.gemm("f32", m, n, k, 1.0, np.abs(A), m, np.abs(B), k, 0.0, bound, m)
    got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B)
    assert "tcgen05" in kern, kern
    assert float(np.max(np.abs(got.astype(np.float64) - want) / np.maximum(bound, 1e-30))) <= 1e-5


@pytest.mark.parametrize("shape", [(128, 128, 8192), (512, 512, 8192), (384, 260, 4100), (200, 130, 2052)],
                         ids=lambda s: "x".join(map(str, s)))
def test_sgemm_tensor_core_split_k(blob, ctx, oracle, shape):
    """Tall-skinny M = N << K: (tile, k-split) work items, partials reduced in split order by
    the last-arriving CTA; repeated calls exercise the self-resetting arrival counters."""
    m, n, k = shape
    A, B, _ = oracle.init_gemm("f32", m, n, k)
    want = oracle.harness_gemm("f32", m, n, k)
    ctx.set_option("sgemm_pair", "off")
    try:
        got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, iters=3)
        got2, _ = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, iters=1)
    finally:
        ctx.set_option("sgemm_pair", "auto")
    assert "splitk" in kern, kern
    assert max_rel_err(got, want) <= 1e-5
    assert np.array_equal(got, got2)  # deterministic reduction order


def test_sgemm_split_k_alpha_beta(blob, ctx, oracle):
    """beta != 0 through the split-K reduction epilogue (applied once, by the last arriver)."""
    rng = np.random.default_rng(6)
    m, n, k = 200, 136, 4104
    lda, ldb, ldc = 200, 4104, 203
    A = rng.random(lda * k).astype(np.float32)
    B = rng.random(ldb * n).astype(np.float32)
    C0 = (rng.random(ldc * n) * 1000).astype(np.float32)
    want = C0.copy()
    oracle.gemm("f32", m, n, k, 0.75, A, lda, B, ldb, 2.0, want, ldc)
    ctx.set_option("sgemm_pair", "off")
    try:
        got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, C0=C0, alpha=0.75, beta=2.0,
                             lda=lda, ldb=ldb, ldc=ldc, iters=1)
    finally:
        ctx.set_option("sgemm_pair", "auto")
    assert "splitk" in kern, kern
    g, w = got.reshape(n, ldc), want.reshape(n, ldc)
    assert np.array_equal(g[:, m:], C0.reshape(n, ldc)[:, m:])  # padding rows untouched
    assert max_rel_err(g[:, :m].ravel(), w[:, :m].ravel()) <= 1e-5


@pytest.mark.parametrize("pair", ["on", "off"])
@pytest.mark.parametrize("shape", [(516, 772, 264), (1028, 260, 72), (2048, 2048, 32), (300, 2304, 72)],
                         ids=lambda s: "x".join(map(str, s)))
def test_sgemm_tma_store_epilogue_matches_register_stores(blob, ctx, oracle, shape, pair):
    """beta == 0 with an aligned C goes through shared memory and cp.async.bulk.tensor stores
    (clipped at the ragged edges by the tensor map); the register-store epilogue must give the
    same bits, and rows / columns outside C must stay untouched."""
    m, n, k = shape
    if pair == "on" and (m < 256 or n < 256):
        pytest.skip("pair tiles need m, n >= 256")
    A, B, _ = oracle.init_gemm("f32", m, n, k)
    ldc = m + 4  # 16-byte aligned, with padding rows the stores must not touch
    C0 = np.full(ldc * n, -7.0, np.float32)
    ctx.set_option("sgemm_pair", pair)
    try:
        got_tma, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, C0=C0, ldc=ldc)
        ctx.set_option("sgemm_tma_store", "off")
        got_reg, kern2 = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, C0=C0, ldc=ldc)
    finally:
        ctx.set_option("sgemm_tma_store", "on")
        ctx.set_option("sgemm_pair", "auto")
    assert "tcgen05" in kern and kern == kern2, (kern, kern2)
    assert np.array_equal(got_tma, got_reg)
    assert np.all(got_tma.reshape(n, ldc)[:, m:] == -7.0)
    want = oracle.harness_gemm("f32", m, n, k)
    assert max_rel_err(got_tma.reshape(n, ldc)[:, :m].ravel(), want) <= 1e-5


# ---- mid-size CTA-pair tile 256 x 128 (cta_group::2, UMMA N = 128), forced ---------------------
PAIR128_SHAPES = [
    (256, 128, 32), (256, 256, 256), (512, 384, 96), (768, 1280, 1000), (260, 300, 200), (1028, 516, 260),
    (1024, 1024, 1024), (1536, 1536, 1536), (300, 2304, 72), (2052, 1028, 516), (256, 128, 32768), (1280, 1152, 2052),
]


@pytest.fixture
def force_pair128(ctx):
    ctx.set_option("sgemm_pair", "128")
    yield
    ctx.set_option("sgemm_pair", "auto")


@pytest.mark.parametrize("shape", PAIR128_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_sgemm_pair128_kernel(blob, ctx, oracle, force_pair128, shape):
    m, n, k = shape
    A, B, _ = oracle.init_gemm("f32", m, n, k)
    want = oracle.harness_gemm("f32", m, n, k)
    got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, iters=2)
    assert "pair_256x128" in kern, kern
    err = max_rel_err(got, want)
    assert err <= 1e-5, (shape, kern, err)


def test_sgemm_pair128_alpha_beta_ld_and_bitwise_vs_single_cta(blob, ctx, oracle, force_pair128):
    rng = np.random.default_rng(14)
    m, n, k = 700, 520, 200
    lda, ldb, ldc = 704, 200, 701
    A = rng.random(lda * k).astype(np.float32)
    B = rng.random(ldb * n).astype(np.float32)
    C0 = rng.random(ldc * n).astype(np.float32)
    want = C0.copy()
    oracle.gemm("f32", m, n, k, 1.5, A, lda, B, ldb, 0.5, want, ldc)
    got, kern = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B, C0=C0, alpha=1.5, beta=0.5,
                         lda=lda, ldb=ldb, ldc=ldc)
    assert "pair_256x128" in kern, kern
    assert max_rel_err(got, want) <= 1e-5
    # same products, same chunked accumulation order as the other tile shapes: identical bits
    m, n, k = 2048, 1536, 264
    A, B, _ = oracle.init_gemm("f32", m, n, k)
    got3, kern3 = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B)
    ctx.set_option("sgemm_pair", "off")
    got1, kern1 = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B)
    ctx.set_option("sgemm_pair", "on")
    got2, kern2 = run_gemm(blob, ctx, "f32", "once", m, n, k, A, B)
    assert "256x128" in kern3 and kern1.endswith("128x128x32") and "256x256" in kern2, (kern1, kern2, kern3)
    assert np.array_equal(got3, got1) and np.array_equal(got3, got2)
