#!/usr/bin/env python
"""Generate the golden fixtures from the REFERENCE's own CPU classes.

Runs in the build container only (needs oracle/_ref/libblobref.so, which is
compiled from /root/reference by oracle/Makefile).  It drives
cpu::gemm_cpu<T> / cpu::gemv_cpu<T> (reference OpenBLAS/gemm.hh:25-43,
OpenBLAS/gemv.hh:25-41) through ::gemm<T>::compute() (include/kernels/gemm.hh:19-42)
and records

  reference_checksums.json  -- the harness checksum (kernels/gemm.hh:61-65,
                               kernels/gemv.hh:61-65) per shape, FP64 and FP32
  reference_outputs.npz     -- the full C / y the reference produced per shape
  reference_samples.npz     -- for shapes whose output is too large to commit (the sizes the
                               TMA-DMMA, tcgen05 single-CTA and CTA-pair kernels take, and the
                               reference's rectangular problem types at D = 8192): 4096 output
                               elements at fixed pseudo-random positions (+ the 4 corners), FP64
                               and FP32, plus their positions

The inputs are not stored: they are the reference's deterministic generator
(srand(19123005), include/kernels/gemm.hh:69-87), which oracle/blob_oracle.c
restates and tests/test_oracle.py pins against the first 16 known values.

    python tests/golden/make_golden.py
"""
import ctypes
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

GEMM_SHAPES = [
    (1, 1, 1), (2, 2, 2), (3, 5, 7), (32, 32, 32), (65, 33, 17), (128, 128, 128),
    (257, 129, 65), (100, 100, 100), (32, 32, 4096), (320, 192, 32), (4, 64, 64),
    (64, 4, 64), (255, 255, 255),
]
# checksum-only (outputs too large to commit): SURVEY.md 9.3 table
GEMM_CHECKSUM_ONLY = [(1024, 1024, 1024), (2048, 2048, 32), (4096, 64, 4096)]
# sampled-element fixtures: square sizes of the flagship kernels + the cfg3 problem types
# (include/doGemm.hh:91-285) at D = 8192
GEMM_SAMPLED = [(1024, 1024, 1024), (2048, 2048, 512), (2048, 2048, 2048), (4096, 4096, 4096),
                (512, 512, 8192), (8192, 512, 512), (512, 8192, 512), (8192, 8192, 512), (8192, 8192, 32),
                (32, 32, 8192), (8192, 32, 32), (32, 8192, 32)]
SAMPLES = 4096


def sample_positions(m, n):
    rng = np.random.default_rng(m * 1000003 + n)
    pos = rng.integers(0, m * n, SAMPLES - 4)
    return np.concatenate([pos, [0, m - 1, m * (n - 1), m * n - 1]]).astype(np.int64)


GEMV_SHAPES = [
    (1, 1), (2, 2), (3, 7), (32, 32), (128, 128), (1024, 1024), (4096, 4096),
    (32, 8192), (8192, 32), (2048, 128), (127, 129), (1001, 77), (33, 2050),
]


def main():
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libblobref.so"))
    lib.blobref_blas_config.restype = ctypes.c_char_p
    dbl_p = ctypes.POINTER(ctypes.c_double)
    sums = {"blas": lib.blobref_blas_config().decode(), "gemm": [], "gemv": []}
    out = {}

    def run_gemm(fn, dtype, m, n, k, keep):
        c = np.zeros(m * n, dtype=dtype)
        cs = ctypes.c_double()
        fn(m, n, k, 1, c.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cs), None)
        if keep:
            out[f"{'d' if dtype == np.float64 else 's'}gemm_{m}_{n}_{k}"] = c
        return cs.value

    def run_gemv(fn, dtype, m, n):
        y = np.zeros(m, dtype=dtype)
        cs = ctypes.c_double()
        fn(m, n, 1, y.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cs), None)
        out[f"{'d' if dtype == np.float64 else 's'}gemv_{m}_{n}"] = y
        return cs.value

    for (m, n, k) in GEMM_SHAPES + GEMM_CHECKSUM_ONLY:
        keep = (m, n, k) in GEMM_SHAPES
        d = run_gemm(lib.blobref_dgemm, np.float64, m, n, k, keep)
        s = run_gemm(lib.blobref_sgemm, np.float32, m, n, k, keep)
        sums["gemm"].append({"m": m, "n": n, "k": k, "dgemm": d, "sgemm": s})
    for (m, n) in GEMV_SHAPES:
        d = run_gemv(lib.blobref_dgemv, np.float64, m, n)
        s = run_gemv(lib.blobref_sgemv, np.float32, m, n)
        sums["gemv"].append({"m": m, "n": n, "dgemv": d, "sgemv": s})

    samples = {}
    for (m, n, k) in GEMM_SAMPLED:
        pos = sample_positions(m, n)
        for fn, dtype, tag in ((lib.blobref_dgemm, np.float64, "d"), (lib.blobref_sgemm, np.float32, "s")):
            c = np.zeros(m * n, dtype=dtype)
            cs = ctypes.c_double()
            fn(m, n, k, 1, c.ctypes.data_as(ctypes.c_void_p), ctypes.byref(cs), None)
            samples[f"{tag}gemm_{m}_{n}_{k}"] = c[pos]
            if not any((r["m"], r["n"], r["k"]) == (m, n, k) for r in sums["gemm"]):
                sums["gemm"].append({"m": m, "n": n, "k": k})
            [r for r in sums["gemm"] if (r["m"], r["n"], r["k"]) == (m, n, k)][0][f"{tag}gemm"] = cs.value
        samples[f"pos_{m}_{n}"] = pos
    np.savez_compressed(os.path.join(HERE, "reference_samples.npz"), **samples)

    with open(os.path.join(HERE, "reference_checksums.json"), "w") as f:
        json.dump(sums, f, indent=1)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote", len(out), "arrays;", len(sums["gemm"]), "gemm +",
          len(sums["gemv"]), "gemv checksums")


if __name__ == "__main__":
    main()
