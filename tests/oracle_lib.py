"""ctypes view of oracle/libbloboracle.so (the C restatement, oracle/blob_oracle.c)
and, when present, oracle/_ref/libblobref.so (the reference's own CPU classes).
Checker-side only: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg -- never by the product package."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "libbloboracle.so")
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libblobref.so")

_vp = ctypes.c_void_p
_i = ctypes.c_int


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.blob_checksum_gemm_f64.restype = ctypes.c_double
        lib.blob_checksum_gemm_f32.restype = ctypes.c_double
        lib.blob_checksum_gemv_f64.restype = ctypes.c_double
        lib.blob_checksum_gemv_f32.restype = ctypes.c_double
        lib.blob_gemm_flops.restype = ctypes.c_ulonglong
        lib.blob_gemv_flops.restype = ctypes.c_ulonglong
        lib.blob_gemm_kib.restype = ctypes.c_double
        lib.blob_gemv_kib.restype = ctypes.c_double
        lib.blob_gemm_f64.argtypes = [_i, _i, _i, _i, _i, ctypes.c_double, _vp, _i, _vp, _i, ctypes.c_double, _vp, _i]
        lib.blob_gemm_f32.argtypes = [_i, _i, _i, _i, _i, ctypes.c_float, _vp, _i, _vp, _i, ctypes.c_float, _vp, _i]
        lib.blob_gemv_f64.argtypes = [_i, _i, _i, ctypes.c_double, _vp, _i, _vp, _i, ctypes.c_double, _vp, _i]
        lib.blob_gemv_f32.argtypes = [_i, _i, _i, ctypes.c_float, _vp, _i, _vp, _i, ctypes.c_float, _vp, _i]

    @staticmethod
    def _np(dtype):
        return np.float64 if dtype == "f64" else np.float32

    def rand_stream(self, count):
        out = np.zeros(count, dtype=np.int32)
        self.lib.blob_rand_stream(count, out.ctypes.data_as(_vp))
        return out

    def init_gemm(self, dtype, m, n, k):
        """A (m*k), B (k*n), C (m*n) exactly as include/kernels/gemm.hh:69-87."""
        t = self._np(dtype)
        A, B, C = np.empty(m * k, t), np.empty(k * n, t), np.empty(m * n, t)
        fn = self.lib.blob_init_gemm_f64 if dtype == "f64" else self.lib.blob_init_gemm_f32
        fn(m, n, k, A.ctypes.data_as(_vp), B.ctypes.data_as(_vp), C.ctypes.data_as(_vp))
        return A, B, C

    def init_gemv(self, dtype, m, n):
        t = self._np(dtype)
        A, x, y = np.empty(m * n, t), np.empty(n, t), np.empty(m, t)
        fn = self.lib.blob_init_gemv_f64 if dtype == "f64" else self.lib.blob_init_gemv_f32
        fn(m, n, A.ctypes.data_as(_vp), x.ctypes.data_as(_vp), y.ctypes.data_as(_vp))
        return A, x, y

    def gemm(self, dtype, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, ta=0, tb=0):
        fn = self.lib.blob_gemm_f64 if dtype == "f64" else self.lib.blob_gemm_f32
        fn(ta, tb, m, n, k, alpha, A.ctypes.data, lda, B.ctypes.data, ldb, beta, C.ctypes.data, ldc)
        return C

    def gemv(self, dtype, m, n, alpha, A, lda, x, incx, beta, y, incy, trans=0):
        fn = self.lib.blob_gemv_f64 if dtype == "f64" else self.lib.blob_gemv_f32
        fn(trans, m, n, alpha, A.ctypes.data, lda, x.ctypes.data, incx, beta, y.ctypes.data, incy)
        return y

    def checksum_gemm(self, dtype, m, n, C):
        fn = self.lib.blob_checksum_gemm_f64 if dtype == "f64" else self.lib.blob_checksum_gemm_f32
        return fn(m, n, C.ctypes.data_as(_vp))

    def checksum_gemv(self, dtype, m, y):
        fn = self.lib.blob_checksum_gemv_f64 if dtype == "f64" else self.lib.blob_checksum_gemv_f32
        return fn(m, y.ctypes.data_as(_vp))

    def harness_gemm(self, dtype, m, n, k):
        """init + C = A*B the way the harness' CPU leg does (alpha=1, beta=0)."""
        A, B, C = self.init_gemm(dtype, m, n, k)
        self.gemm(dtype, m, n, k, 1.0, A, max(1, m), B, max(1, k), 0.0, C, max(1, m))
        return C

    def harness_gemv(self, dtype, m, n):
        A, x, y = self.init_gemv(dtype, m, n)
        self.gemv(dtype, m, n, 1.0, A, max(1, m), x, 1, 0.0, y, 1)
        return y


def load_oracle():
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    return Oracle(ctypes.CDLL(ORACLE_SO))


def load_reference():
    """The reference's own classes (OpenBLAS); None when not built."""
    if not os.path.exists(REF_SO):
        return None
    lib = ctypes.CDLL(REF_SO)
    lib.blobref_blas_config.restype = ctypes.c_char_p
    return lib


def max_rel_err(got, want):
    """Element-wise max relative error (north_star: <= 1e-12 FP64, <= 1e-5 FP32)."""
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    denom = np.maximum(np.abs(want), np.finfo(np.float64).tiny)
    return float(np.max(np.abs(got - want) / denom)) if got.size else 0.0
