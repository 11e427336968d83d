"""gpu-blas-offload-benchmark_b200 -- the B200-native (sm_100a) GPU BLAS backend
for GPU-BLOB.

The product is native: `lib/libb200blas.so` (hand-written CUDA behind the C ABI
declared in `include/b200blas.h`) plus the C++ backend classes in `B200/`
(gpu::gemm_gpu<T>, gpu::gemv_gpu<T>) that plug into the unmodified GPU-BLOB
harness.  This Python package is only a ctypes view of that ABI so that
`tests/` and `bench.py` can call exactly what the harness calls.

There is no CPU path here: loading fails loudly if the CUDA library has not
been built, and every entry point fails if there is no B200.

The directory name contains hyphens; import it with
    importlib.import_module("gpu-blas-offload-benchmark_b200")
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "lib", "libb200blas.so")
HEADER_PATH = os.path.join(os.path.dirname(_PKG), "include", "b200blas.h")

OFFLOAD_ALWAYS, OFFLOAD_ONCE, OFFLOAD_UNIFIED = 0, 1, 2  # include/utilities.hh:50-54
OP_N, OP_T = 0, 1
NUM_LANES = 3


class B200Error(RuntimeError):
    pass


_c_int, _c_size_t, _c_void_p = ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p
_c_float, _c_double = ctypes.c_float, ctypes.c_double
_pp = ctypes.POINTER(ctypes.c_void_p)

# name -> (restype, argtypes); mirrors include/b200blas.h one to one
SIGNATURES = {
    "b200_ctx_create": (_c_int, [_pp, _c_int]),
    "b200_ctx_destroy": (_c_int, [_c_void_p]),
    "b200_ctx_set_stream": (_c_int, [_c_void_p, _c_void_p]),
    "b200_ctx_device": (_c_int, [_c_void_p, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int)]),
    "b200_ctx_set_option": (_c_int, [_c_void_p, ctypes.c_char_p, ctypes.c_char_p]),
    "b200_alloc": (_c_int, [_c_void_p, _c_int, _c_size_t, _pp, _pp]),
    "b200_free": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p]),
    "b200_h2d_async": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int]),
    "b200_d2h_async": (_c_int, [_c_void_p, _c_void_p, _c_void_p, _c_size_t, _c_int]),
    "b200_prefetch_async": (_c_int, [_c_void_p, _c_void_p, _c_size_t, _c_int, _c_int]),
    "b200_advise": (_c_int, [_c_void_p, _c_void_p, _c_size_t, _c_int]),
    "b200_sync": (_c_int, [_c_void_p]),
    "b200_sgemm": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_float, _c_void_p,
                            _c_int, _c_void_p, _c_int, _c_float, _c_void_p, _c_int]),
    "b200_dgemm": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_double, _c_void_p,
                            _c_int, _c_void_p, _c_int, _c_double, _c_void_p, _c_int]),
    "b200_sgemv": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p, _c_int,
                            _c_void_p, _c_int, _c_float, _c_void_p, _c_int]),
    "b200_dgemv": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_double, _c_void_p, _c_int,
                            _c_void_p, _c_int, _c_double, _c_void_p, _c_int]),
    "b200_sgemm_offload": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_float, _c_void_p, _c_void_p,
                                    _c_int, _c_void_p, _c_void_p, _c_int, _c_float, _c_void_p,
                                    _c_void_p, _c_int]),
    "b200_dgemm_offload": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_double, _c_void_p, _c_void_p,
                                    _c_int, _c_void_p, _c_void_p, _c_int, _c_double, _c_void_p,
                                    _c_void_p, _c_int]),
    "b200_sgemv_offload": (_c_int, [_c_void_p, _c_int, _c_int, _c_float, _c_void_p, _c_void_p, _c_int,
                                    _c_void_p, _c_void_p, _c_float, _c_void_p, _c_void_p]),
    "b200_dgemv_offload": (_c_int, [_c_void_p, _c_int, _c_int, _c_double, _c_void_p, _c_void_p, _c_int,
                                    _c_void_p, _c_void_p, _c_double, _c_void_p, _c_void_p]),
    "b200_last_kernel": (ctypes.c_char_p, [_c_void_p]),
    "b200_launch_count": (ctypes.c_ulonglong, [_c_void_p]),
    "b200_timer_start": (_c_int, [_c_void_p]),
    "b200_timer_stop": (_c_int, [_c_void_p, ctypes.POINTER(_c_float)]),
    "b200_probe_peak": (_c_int, [_c_void_p, ctypes.c_char_p, ctypes.POINTER(_c_double)]),
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_version": (ctypes.c_char_p, []),
}

_lib = None


def load():
    """dlopen libb200blas.so and bind every symbol of the header.  Loud on failure."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise B200Error(
            f"{LIB_PATH} is missing: build it with `python gpu-blas-offload-benchmark_b200/build.py` "
            "(or __graft_entry__.build()).  There is no CPU fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def _check(rc):
    if rc != 0:
        raise B200Error(load().b200_last_error().decode())


def _ptr(x):
    """Raw address of a torch tensor / numpy array / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    if hasattr(x, "ctypes"):
        return x.ctypes.data
    raise TypeError(type(x))


class Context:
    """RAII view of a b200_ctx (the cublasHandle + 3 streams of the reference,
    cuBLAS/gemm.hh:46-63)."""

    def __init__(self, device=-1):
        self._lib = load()
        h = ctypes.c_void_p()
        _check(self._lib.b200_ctx_create(ctypes.byref(h), device))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200_ctx_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing
    def set_stream(self, cuda_stream):
        _check(self._lib.b200_ctx_set_stream(self._h, cuda_stream))

    def set_option(self, key, value):
        _check(self._lib.b200_ctx_set_option(self._h, key.encode(), str(value).encode()))

    def device(self):
        d, s = ctypes.c_int(), ctypes.c_int()
        _check(self._lib.b200_ctx_device(self._h, ctypes.byref(d), ctypes.byref(s)))
        return d.value, s.value

    def alloc(self, mode, nbytes):
        h, d = ctypes.c_void_p(), ctypes.c_void_p()
        _check(self._lib.b200_alloc(self._h, mode, nbytes, ctypes.byref(h), ctypes.byref(d)))
        return h.value, d.value

    def free(self, mode, host, dev):
        _check(self._lib.b200_free(self._h, mode, host, dev))

    def h2d(self, dev, host, nbytes, lane=0):
        _check(self._lib.b200_h2d_async(self._h, _ptr(dev), _ptr(host), nbytes, lane))

    def d2h(self, host, dev, nbytes, lane=2):
        _check(self._lib.b200_d2h_async(self._h, _ptr(host), _ptr(dev), nbytes, lane))

    def prefetch(self, ptr, nbytes, to_device, lane=0):
        _check(self._lib.b200_prefetch_async(self._h, _ptr(ptr), nbytes, int(to_device), lane))

    def advise(self, ptr, nbytes, is_input):
        _check(self._lib.b200_advise(self._h, _ptr(ptr), nbytes, int(is_input)))

    def sync(self):
        _check(self._lib.b200_sync(self._h))

    # -- compute (device / managed addresses)
    def gemm(self, dtype, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, transa=OP_N, transb=OP_N):
        fn = self._lib.b200_sgemm if dtype == "f32" else self._lib.b200_dgemm
        _check(fn(self._h, transa, transb, m, n, k, alpha, _ptr(A), lda, _ptr(B), ldb, beta, _ptr(C), ldc))

    def gemv(self, dtype, m, n, alpha, A, lda, x, incx, beta, y, incy, trans=OP_N):
        fn = self._lib.b200_sgemv if dtype == "f32" else self._lib.b200_dgemv
        _check(fn(self._h, trans, m, n, alpha, _ptr(A), lda, _ptr(x), incx, beta, _ptr(y), incy))

    def gemm_offload(self, dtype, m, n, k, alpha, hA, dA, lda, hB, dB, ldb, beta, hC, dC, ldc):
        fn = self._lib.b200_sgemm_offload if dtype == "f32" else self._lib.b200_dgemm_offload
        _check(fn(self._h, m, n, k, alpha, _ptr(hA), _ptr(dA), lda, _ptr(hB), _ptr(dB), ldb, beta,
                  _ptr(hC), _ptr(dC), ldc))

    def gemv_offload(self, dtype, m, n, alpha, hA, dA, lda, hx, dx, beta, hy, dy):
        fn = self._lib.b200_sgemv_offload if dtype == "f32" else self._lib.b200_dgemv_offload
        _check(fn(self._h, m, n, alpha, _ptr(hA), _ptr(dA), lda, _ptr(hx), _ptr(dx), beta, _ptr(hy), _ptr(dy)))

    # -- introspection
    def last_kernel(self):
        return self._lib.b200_last_kernel(self._h).decode()

    def launch_count(self):
        return int(self._lib.b200_launch_count(self._h))

    def timer_start(self):
        _check(self._lib.b200_timer_start(self._h))

    def timer_stop(self):
        ms = ctypes.c_float()
        _check(self._lib.b200_timer_stop(self._h, ctypes.byref(ms)))
        return ms.value

    def probe_peak(self, which):
        r = ctypes.c_double()
        _check(self._lib.b200_probe_peak(self._h, which.encode(), ctypes.byref(r)))
        return r.value


def host_view(ptr, count, dtype):
    """numpy view of `count` elements of pinned/managed host memory at `ptr`."""
    import numpy as np
    ct = {"f32": ctypes.c_float, "f64": ctypes.c_double}[dtype]
    buf = (ct * count).from_address(ptr)
    return np.frombuffer(buf, dtype=np.float32 if dtype == "f32" else np.float64, count=count)
