"""GPU parity tests proper: the hand-written CUDA path, called through the C ABI
(include/b200blas.h) the way B200/gemm.hh and B200/gemv.hh call it, compared
with the oracle (oracle/blob_oracle.c) on the reference's own deterministic
inputs and with the committed reference outputs (tests/golden/).

Tolerances are BASELINE.json's: element-wise max relative error <= 1e-12 (FP64)
and <= 1e-5 (FP32), plus the harness checksum within the same bound.
"""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle_lib import max_rel_err

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_checksums.json")))
OUTS = np.load(os.path.join(HERE, "golden", "reference_outputs.npz"))
TOL = {"f64": 1e-12, "f32": 1e-5}
NP = {"f64": np.float64, "f32": np.float32}
MODES = {"always": 0, "once": 1, "unified": 2}


class Operand:
    """pinned+device (or managed) buffer obtained through b200_alloc."""

    def __init__(self, blob, ctx, mode, dtype, count):
        self.blob, self.ctx, self.mode, self.dtype, self.count = blob, ctx, mode, dtype, count
        self.nbytes = max(1, count) * np.dtype(NP[dtype]).itemsize
        self.host, self.dev = ctx.alloc(mode, self.nbytes)
        self.np = blob.host_view(self.host, max(1, count), dtype)[:count]

    def free(self):
        self.ctx.free(self.mode, self.host, self.dev)


def run_gemm(blob, ctx, dtype, mode, m, n, k, A, B, C0=None, alpha=1.0, beta=0.0, iters=1,
             lda=None, ldb=None, ldc=None, ta=0, tb=0):
    """One harness-style compute(): preLoop + iters x call + postLoop, in `mode`."""
    md = MODES[mode]
    lda = lda or max(1, k if ta else m)
    ldb = ldb or max(1, n if tb else k)
    ldc = ldc or max(1, m)
    a = Operand(blob, ctx, md, dtype, A.size)
    b = Operand(blob, ctx, md, dtype, B.size)
    c = Operand(blob, ctx, md, dtype, ldc * n)
    try:
        a.np[:] = A
        b.np[:] = B
        c.np[:] = 0 if C0 is None else C0
        if mode == "once":
            ctx.h2d(a.dev, a.host, a.nbytes, 0)
            ctx.h2d(b.dev, b.host, b.nbytes, 1)
            ctx.h2d(c.dev, c.host, c.nbytes, 2)
        elif mode == "unified":
            ctx.prefetch(a.dev, a.nbytes, True, 0)
            ctx.prefetch(b.dev, b.nbytes, True, 1)
            ctx.prefetch(c.dev, c.nbytes, True, 2)
        for _ in range(iters):
            if mode == "always":
                assert not ta and not tb
                ctx.gemm_offload(dtype, m, n, k, alpha, a.host, a.dev, lda, b.host, b.dev, ldb, beta,
                                 c.host, c.dev, ldc)
            else:
                ctx.gemm(dtype, m, n, k, alpha, a.dev, lda, b.dev, ldb, beta, c.dev, ldc, transa=ta, transb=tb)
        if mode == "once":
            ctx.d2h(c.host, c.dev, c.nbytes, 2)
        elif mode == "unified":
            ctx.prefetch(c.dev, c.nbytes, False, 2)
        ctx.sync()
        return c.np.copy(), ctx.last_kernel()
    finally:
        a.free(); b.free(); c.free()


def run_gemv(blob, ctx, dtype, mode, m, n, A, x, y0=None, alpha=1.0, beta=0.0, iters=1, lda=None,
             trans=0, incx=1, incy=1):
    md = MODES[mode]
    lda = lda or max(1, m)
    leny = n if trans else m
    a = Operand(blob, ctx, md, dtype, A.size)
    xv = Operand(blob, ctx, md, dtype, x.size)
    yv = Operand(blob, ctx, md, dtype, max(1, (leny - 1) * abs(incy) + 1) if leny else 0)
    try:
        a.np[:] = A
        xv.np[:] = x
        yv.np[:] = 0 if y0 is None else y0
        if mode == "once":
            ctx.h2d(a.dev, a.host, a.nbytes, 0)
            ctx.h2d(xv.dev, xv.host, xv.nbytes, 1)
            ctx.h2d(yv.dev, yv.host, yv.nbytes, 2)
        elif mode == "unified":
            ctx.prefetch(a.dev, a.nbytes, True, 0)
            ctx.prefetch(xv.dev, xv.nbytes, True, 1)
            ctx.prefetch(yv.dev, yv.nbytes, True, 2)
        for _ in range(iters):
            if mode == "always":
                assert not trans and incx == 1 and incy == 1
                ctx.gemv_offload(dtype, m, n, alpha, a.host, a.dev, lda, xv.host, xv.dev, beta, yv.host, yv.dev)
            else:
                ctx.gemv(dtype, m, n, alpha, a.dev, lda, xv.dev, incx, beta, yv.dev, incy, trans=trans)
        if mode == "once":
            ctx.d2h(yv.host, yv.dev, yv.nbytes, 2)
        elif mode == "unified":
            ctx.prefetch(yv.dev, yv.nbytes, False, 2)
        ctx.sync()
        return yv.np.copy(), ctx.last_kernel()
    finally:
        a.free(); xv.free(); yv.free()


# ------------------------------------------------------------------ golden ---

@pytest.mark.parametrize("mode", ["once", "always", "unified"])
@pytest.mark.parametrize("name", sorted(OUTS.files))
def test_matches_reference_outputs(blob, ctx, oracle, name, mode):
    """Same inputs as the reference's CPU leg -> its committed OpenBLAS outputs."""
    kind, *dims = name.split("_")
    dims = [int(d) for d in dims]
    dt = "f64" if kind[0] == "d" else "f32"
    if kind.endswith("gemm"):
        m, n, k = dims
        A, B, _ = oracle.init_gemm(dt, m, n, k)
        got, kern = run_gemm(blob, ctx, dt, mode, m, n, k, A, B, iters=2)
        cs = oracle.checksum_gemm(dt, m, n, got)
        gold = [r for r in GOLD["gemm"] if (r["m"], r["n"], r["k"]) == (m, n, k)][0][kind]
    else:
        m, n = dims
        A, x, _ = oracle.init_gemv(dt, m, n)
        got, kern = run_gemv(blob, ctx, dt, mode, m, n, A, x, iters=2)
        cs = oracle.checksum_gemv(dt, m, got)
        gold = [r for r in GOLD["gemv"] if (r["m"], r["n"]) == (m, n)][0][kind]
    err = max_rel_err(got, OUTS[name])
    assert err <= TOL[dt], (name, mode, kern, err)
    assert abs(cs - gold) <= TOL[dt] * abs(gold), (name, mode, kern, cs, gold)


@pytest.mark.parametrize("row", [r for r in GOLD["gemm"] if r["m"] * r["n"] * r["k"] > 5e8],
                         ids=lambda r: f"{r['m']}x{r['n']}x{r['k']}")
def test_large_golden_checksums(blob, ctx, oracle, row):
    """SURVEY.md 9.3 rows whose outputs are too large to commit: checksum only."""
    m, n, k = row["m"], row["n"], row["k"]
    for dt, key in (("f64", "dgemm"), ("f32", "sgemm")):
        A, B, _ = oracle.init_gemm(dt, m, n, k)
        got, kern = run_gemm(blob, ctx, dt, "once", m, n, k, A, B)
        cs = oracle.checksum_gemm(dt, m, n, got)
        assert abs(cs - row[key]) <= TOL[dt] * abs(row[key]), (dt, kern, cs, row[key])


SAMPLES = np.load(os.path.join(HERE, "golden", "reference_samples.npz"))
# which kernel family must have produced each sampled shape (so the fixture really meets the
# flagship kernels, not the small-tile fallbacks)
SAMPLED_KERNELS = {
    ("f64", (2048, 2048, 2048)): "dgemm_dmma_tma", ("f64", (4096, 4096, 4096)): "dgemm_dmma_tma",
    ("f64", (8192, 8192, 512)): "dgemm_dmma_tma", ("f64", (512, 512, 8192)): "dgemm_dmma_tma",
    ("f32", (2048, 2048, 2048)): "sgemm_3xtf32_tcgen05", ("f32", (4096, 4096, 4096)): "sgemm_3xtf32_tcgen05_pair",
    ("f32", (8192, 8192, 512)): "sgemm_3xtf32_tcgen05_pair", ("f32", (1024, 1024, 1024)): "sgemm_3xtf32_tcgen05",
    ("f32", (512, 512, 8192)): "sgemm_3xtf32_tcgen05",
}


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("shape", sorted({tuple(int(x) for x in k.split("_")[1:]) for k in SAMPLES.files if k[1:5] == "gemm"}),
                         ids=lambda s: "x".join(map(str, s)))
def test_sampled_reference_outputs(blob, ctx, oracle, shape, dt):
    """4096 output elements per shape as the REFERENCE's own OpenBLAS classes produced them
    (tests/golden/make_golden.py), at the sizes the TMA-DMMA, tcgen05 and CTA-pair kernels take
    and at the reference's rectangular problem types for D = 8192 (include/doGemm.hh:91-285)."""
    m, n, k = shape
    A, B, _ = oracle.init_gemm(dt, m, n, k)
    got, kern = run_gemm(blob, ctx, dt, "once", m, n, k, A, B)
    pos = SAMPLES[f"pos_{m}_{n}"]
    want = SAMPLES[f"{'d' if dt == 'f64' else 's'}gemm_{m}_{n}_{k}"]
    err = max_rel_err(got[pos], want)
    assert err <= TOL[dt], (dt, shape, kern, err)
    must = SAMPLED_KERNELS.get((dt, shape))
    if must:
        assert kern.replace("_fused", "").startswith(must), (dt, shape, kern)


# ------------------------------------------------------- oracle, many shapes ---

GEMM_SHAPES = [
    (1, 1, 1), (1, 7, 3), (7, 1, 3), (5, 5, 1), (17, 19, 23), (31, 33, 35), (63, 65, 64), (64, 64, 64),
    (96, 96, 96), (127, 129, 131), (128, 128, 128), (129, 127, 16), (200, 120, 50), (256, 256, 256),
    (257, 255, 129), (384, 512, 96), (513, 515, 517), (640, 384, 1000), (1024, 1024, 64),
    (32, 32, 2048), (32, 32, 8191), (48, 40, 4096), (1000, 32, 32), (32, 1000, 32), (1024, 32, 32),
    (2048, 2048, 32), (33, 2047, 65), (1536, 130, 700),
]


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("shape", GEMM_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_gemm_vs_oracle(blob, ctx, oracle, shape, dt):
    m, n, k = shape
    A, B, _ = oracle.init_gemm(dt, m, n, k)
    want = oracle.harness_gemm(dt, m, n, k)
    got, kern = run_gemm(blob, ctx, dt, "once", m, n, k, A, B)
    err = max_rel_err(got, want)
    assert err <= TOL[dt], (shape, dt, kern, err)


GEMV_SHAPES = [
    (1, 1), (1, 9), (9, 1), (31, 33), (64, 64), (100, 3000), (127, 129), (255, 1025), (1000, 1000),
    (1024, 1024), (2049, 2047), (4096, 512), (512, 4096), (32, 32768), (32768, 32), (8192, 2048),
    (2048, 8192), (16384, 16), (6001, 3001), (4098, 1026),
]


@pytest.mark.parametrize("dt", ["f64", "f32"])
@pytest.mark.parametrize("shape", GEMV_SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_gemv_vs_oracle(blob, ctx, oracle, shape, dt):
    m, n = shape
    A, x, _ = oracle.init_gemv(dt, m, n)
    want = oracle.harness_gemv(dt, m, n)
    got, kern = run_gemv(blob, ctx, dt, "once", m, n, A, x)
    err = max_rel_err(got, want)
    assert err <= TOL[dt], (shape, dt, kern, err)


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_blas_generality(blob, ctx, oracle, dt):
    """Transposes, alpha/beta, padded leading dimensions, increments:
    parity unpinned by the reference, oracle = the BLAS definition."""
    rng = np.random.default_rng(11)
    t = NP[dt]
    m, n, k = 150, 70, 90
    for ta in (0, 1):
        for tb in (0, 1):
            lda = (k if ta else m) + 3
            ldb = (n if tb else k) + 5
            ldc = m + 2
            A = rng.random(lda * (m if ta else k)).astype(t)
            B = rng.random(ldb * (k if tb else n)).astype(t)
            C0 = rng.random(ldc * n).astype(t)
            want = C0.copy()
            oracle.gemm(dt, m, n, k, 1.25, A, lda, B, ldb, -0.75, want, ldc, ta=ta, tb=tb)
            got, kern = run_gemm(blob, ctx, dt, "once", m, n, k, A, B, C0=C0, alpha=1.25, beta=-0.75,
                                 lda=lda, ldb=ldb, ldc=ldc, ta=ta, tb=tb)
            # padding rows of C must be untouched (exact), the rest within tolerance
            assert max_rel_err(got, want) <= TOL[dt], (ta, tb, kern)
    m, n = 300, 170
    lda = m + 1
    A = rng.random(lda * n).astype(t)
    for trans in (0, 1):
        lenx, leny = (m, n) if trans else (n, m)
        for incx, incy in ((1, 1), (2, 3)):
            x = rng.random((lenx - 1) * incx + 1).astype(t)
            y0 = rng.random((leny - 1) * incy + 1).astype(t)
            want = y0.copy()
            oracle.gemv(dt, m, n, 0.5, A, lda, x, incx, 2.0, want, incy, trans=trans)
            got, kern = run_gemv(blob, ctx, dt, "once", m, n, A, x, y0=y0, alpha=0.5, beta=2.0, lda=lda,
                                 trans=trans, incx=incx, incy=incy)
            assert max_rel_err(got, want) <= TOL[dt], (trans, incx, incy, kern)


def test_degenerate_dimensions(blob, ctx):
    """m, n or k == 0 and invalid arguments follow BLAS: quick return / error code."""
    c = ctx
    h, d = c.alloc(1, 4096)
    try:
        c.gemm("f64", 0, 5, 5, 1.0, d, 1, d, 5, 0.0, d, 1)
        c.gemm("f64", 5, 0, 5, 1.0, d, 5, d, 5, 0.0, d, 5)
        c.gemv("f32", 0, 5, 1.0, d, 1, d, 1, 0.0, d, 1)
        c.sync()
        with pytest.raises(blob.B200Error, match="lda"):
            c.gemm("f64", 8, 8, 8, 1.0, d, 4, d, 8, 0.0, d, 8)
        with pytest.raises(blob.B200Error, match="negative"):
            c.gemv("f64", -1, 8, 1.0, d, 1, d, 1, 0.0, d, 1)
        # k == 0 with beta == 0 must zero C
        v = blob.host_view(h, 16, "f64")
        v[:] = 7.0
        c.h2d(d, h, 128, 0)
        c.gemm("f64", 4, 4, 0, 1.0, d, 4, d, 1, 0.0, d, 4)
        c.d2h(h, d, 128, 2)
        c.sync()
        assert not v.any()
    finally:
        c.free(1, h, d)


def test_results_are_deterministic(blob, ctx, oracle):
    """Split-K / split-N reductions run in a fixed order: bit-identical reruns."""
    for dt in ("f64", "f32"):
        A, B, _ = oracle.init_gemm(dt, 32, 32, 4096)
        r1, _ = run_gemm(blob, ctx, dt, "once", 32, 32, 4096, A, B)
        r2, _ = run_gemm(blob, ctx, dt, "once", 32, 32, 4096, A, B)
        assert np.array_equal(r1, r2)
        A, x, _ = oracle.init_gemv(dt, 32, 8192)
        y1, _ = run_gemv(blob, ctx, dt, "once", 32, 8192, A, x)
        y2, _ = run_gemv(blob, ctx, dt, "once", 32, 8192, A, x)
        assert np.array_equal(y1, y2)


# ------------------------------------------ full BASELINE sizes: properties ---

def _device_fill(torch, shape, dtype, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    # values k/7 with k in [0, 100): the reference's value distribution
    return (torch.randint(0, 100, shape, generator=g, device="cuda").to(dtype) / 7.0)


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_full_size_gemm_blockwise(blob, ctx, oracle, dt):
    """8192^3 (BASELINE config 2): C blocks picked from corners/middle equal the
    oracle's product of the matching A row-panel and B column-panel."""
    import torch
    n = 8192
    tdt = torch.float64 if dt == "f64" else torch.float32
    A = _device_fill(torch, (n * n,), tdt, 1)
    B = _device_fill(torch, (n * n,), tdt, 2) * (7.0 / 3.0)
    C = torch.empty(n * n, dtype=tdt, device="cuda")
    torch.cuda.synchronize()
    ctx.gemm(dt, n, n, n, 1.0, A, n, B, n, 0.0, C, n)
    ctx.sync()
    kern = ctx.last_kernel()
    Ah = A.view(n, n).cpu().numpy()  # [k][m] storage: column-major A(m,k) = Ah[k, m]
    Bh = B.view(n, n).cpu().numpy()  # column-major B(k,j) = Bh[j, k]
    Ch = C.view(n, n).cpu().numpy()  # C(i,j) = Ch[j, i]
    bs = 48
    for (i0, j0) in ((0, 0), (n - bs, n - bs), (0, n - bs), (4000, 4100), (8100, 17)):
        Ap = np.ascontiguousarray(Ah[:, i0:i0 + bs])           # k x bs  -> A panel as column-major (bs x k), lda = bs
        a_cm = np.ascontiguousarray(Ap).reshape(-1)             # element (i,p) at p*bs + i
        Bp = np.ascontiguousarray(Bh[j0:j0 + bs, :]).reshape(-1)  # element (p,j) at j*n + p
        want = np.zeros(bs * bs, dtype=NP[dt])
        oracle.gemm(dt, bs, bs, n, 1.0, a_cm, bs, Bp, n, 0.0, want, bs)
        got = np.ascontiguousarray(Ch[j0:j0 + bs, i0:i0 + bs]).reshape(-1)
        err = max_rel_err(got, want)
        assert err <= TOL[dt], (dt, kern, (i0, j0), err)


@pytest.mark.parametrize("dt", ["f64", "f32"])
def test_full_size_gemv_linearity_and_rows(blob, ctx, oracle, dt):
    """32768^2 (BASELINE config 4): sampled rows equal the oracle's dot products,
    and A(x1 + x2) == A x1 + A x2 within tolerance."""
    import torch
    m = n = 32768
    tdt = torch.float64 if dt == "f64" else torch.float32
    A = _device_fill(torch, (m * n,), tdt, 3)
    x1 = _device_fill(torch, (n,), tdt, 4) * (7.0 / 3.0)
    x2 = _device_fill(torch, (n,), tdt, 5) * (7.0 / 3.0)
    ys = [torch.empty(m, dtype=tdt, device="cuda") for _ in range(3)]
    torch.cuda.synchronize()
    for xv, yv in zip((x1, x2, x1 + x2), ys):
        ctx.gemv(dt, m, n, 1.0, A, m, xv, 1, 0.0, yv, 1)
    ctx.sync()
    kern = ctx.last_kernel()
    y1, y2, y12 = (y.cpu().numpy() for y in ys)
    assert max_rel_err(y12, y1.astype(np.float64) + y2) <= 4 * TOL[dt], kern
    rows = [0, 1, 4097, m // 2, m - 1]
    Arows = A.view(n, m)[:, rows].cpu().numpy().astype(np.float64)  # n x len(rows)
    want = Arows.T @ x1.cpu().numpy().astype(np.float64)
    assert max_rel_err(y1[rows], want) <= TOL[dt], kern


# ---------------------------------------------- C++ backend classes + harness ---

PARITY = os.path.join(ROOT, "gpu-blas-offload-benchmark_b200", "lib", "b200_parity")
BLOB = os.path.join(ROOT, "gpu-blas-offload-benchmark_b200", "lib", "gpu-blob-b200")


@pytest.mark.parametrize("mode", ["once", "always", "unified"])
@pytest.mark.parametrize("kern,dims", [("dgemm", (257, 129, 65)), ("sgemm", (257, 129, 65)),
                                       ("dgemm", (512, 512, 512)), ("sgemm", (1024, 256, 384)),
                                       ("dgemv", (2048, 128)), ("sgemv", (1001, 77)),
                                       ("dgemv", (4096, 4096)), ("sgemv", (32, 8192))])
def test_cpp_backend_classes(oracle, tmp_path, kern, dims, mode):
    """gpu::gemm_gpu<T> / gpu::gemv_gpu<T> driven like include/doGemm.hh:326-340."""
    if not os.path.exists(PARITY):
        pytest.skip("b200_parity not built")
    out = tmp_path / "out.bin"
    dt = "f64" if kern[0] == "d" else "f32"
    args = [PARITY, kern, mode] + [str(d) for d in dims] + ["3", str(out)]
    r = subprocess.run(args, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    checksum = float(r.stdout.split()[0])
    got = np.fromfile(out, dtype=NP[dt])
    if kern.endswith("gemm"):
        want = oracle.harness_gemm(dt, *dims)
        cs = oracle.checksum_gemm(dt, dims[0], dims[1], want)
    else:
        want = oracle.harness_gemv(dt, *dims)
        cs = oracle.checksum_gemv(dt, dims[0], want)
    assert max_rel_err(got, want) <= TOL[dt]
    assert abs(checksum - cs) <= TOL[dt] * abs(cs)


def test_unmodified_harness_runs_end_to_end(tmp_path):
    """The reference harness (doGemm.hh / doGemv.hh, CSV writer, checksum gate,
    threshold tables) with GPU_LIB=B200: CPU + 3 GPU modes, dims 1..40."""
    if not os.path.exists(BLOB):
        pytest.skip("gpu-blob-b200 not built")
    out = tmp_path / "csv"
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run([BLOB, "-i", "3", "-s", "1", "-d", "40", "-o", str(out)], capture_output=True,
                       text=True, timeout=900, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "checksums do not match" not in r.stderr
    files = sorted(os.listdir(out))
    assert len(files) == 28, files
    rows = open(out / "dgemm_square_square_M=N=K.csv").read().strip().splitlines()
    assert rows[0].startswith("Device,Kernel,M,N,K")
    devices = [l.split(",")[0] for l in rows[1:5]]
    assert devices == ["cpu", "gpu_offloadOnce", "gpu_offloadAlways", "gpu_unified"]
