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
_c_ll = ctypes.c_longlong

MG_GEMM, MG_GEMV = 0, 1
MG_UPLOAD, MG_REPLICATE, MG_COMPUTE, MG_DOWNLOAD, MG_GATHER = 1, 2, 4, 8, 16
MG_HANDLE_BYTES = 64


class MgGemmArgs(ctypes.Structure):  # b200_mg_gemm_args
    _fields_ = [("elem", _c_int), ("active", _c_int), ("m", _c_int), ("n_slab", _c_int), ("k", _c_int),
                ("alpha", _c_double), ("beta", _c_double), ("A_rep", _pp), ("lda", _c_int),
                ("hA_share", _c_void_p), ("hB", _c_void_p), ("dB", _c_void_p), ("ldb", _c_int),
                ("hC", _c_void_p), ("dC", _c_void_p), ("ldc", _c_int), ("C_gather", _c_void_p),
                ("phases", _c_int)]


class MgGemvArgs(ctypes.Structure):  # b200_mg_gemv_args
    _fields_ = [("elem", _c_int), ("active", _c_int), ("m_blk", _c_int), ("n", _c_int),
                ("alpha", _c_double), ("beta", _c_double), ("dA", _c_void_p), ("lda", _c_int),
                ("hA", _c_void_p), ("h_lda", _c_ll), ("x_rep", _pp), ("hx_share", _c_void_p),
                ("hy", _c_void_p), ("dy", _c_void_p), ("y_gather", _c_void_p), ("phases", _c_int)]

# name -> (restype, argtypes); mirrors include/b200blas.h one to one
SIGNATURES = {
    "b200_ctx_create": (_c_int, [_pp, _c_int]),
    "b200_ctx_destroy": (_c_int, [_c_void_p]),
    "b200_ctx_set_stream": (_c_int, [_c_void_p, _c_void_p]),
    "b200_ctx_device": (_c_int, [_c_void_p, ctypes.POINTER(_c_int), ctypes.POINTER(_c_int)]),
    "b200_ctx_set_option": (_c_int, [_c_void_p, ctypes.c_char_p, ctypes.c_char_p]),
    "b200_alloc": (_c_int, [_c_void_p, _c_int, _c_size_t, _pp, _pp]),
    "b200_free": (_c_int, [_c_void_p, _c_int, _c_void_p, _c_void_p]),
    "b200_alloc_device": (_c_int, [_c_void_p, _c_size_t, _pp]),
    "b200_free_device": (_c_int, [_c_void_p, _c_void_p]),
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
    "b200_sgemm_prepare": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int]),
    # multi-GPU teams
    "b200_mg_partition": (_c_int, [_c_ll, _c_int, _c_int, _c_int, ctypes.POINTER(_c_ll), ctypes.POINTER(_c_ll)]),
    "b200_mg_active_gemm": (_c_int, [_c_int, _c_int, _c_int, _c_int, _c_int]),
    "b200_mg_active_gemv": (_c_int, [_c_int, _c_int, _c_int, _c_int]),
    "b200_mg_create": (_c_int, [_pp, _c_int]),
    "b200_mg_create_rank": (_c_int, [_pp, _c_int, _c_int, _c_int]),
    "b200_mg_destroy": (_c_int, [_c_void_p]),
    "b200_mg_size": (_c_int, [_c_void_p]),
    "b200_mg_ctx": (_c_int, [_c_void_p, _c_int, _pp]),
    "b200_mg_sync": (_c_int, [_c_void_p]),
    "b200_mg_export": (_c_int, [_c_void_p, _c_void_p, _c_void_p]),
    "b200_mg_import": (_c_int, [_c_void_p, _c_int, _c_void_p, _pp]),
    "b200_mg_unmap": (_c_int, [_c_void_p, _c_void_p]),
    "b200_mg_gemm": (_c_int, [_c_void_p, _c_int, _c_void_p]),
    "b200_mg_gemv": (_c_int, [_c_void_p, _c_int, _c_void_p]),
    "b200_mg_gemm_team": (_c_int, [_c_void_p, _c_void_p]),
    "b200_mg_gemv_team": (_c_int, [_c_void_p, _c_void_p]),
    "b200_mg_problem_create": (_c_int, [_c_void_p, _c_int, _c_int, _c_int, _c_int, _c_int, _c_int, _pp, _pp, _pp, _pp]),
    "b200_mg_problem_active": (_c_int, [_c_void_p]),
    "b200_mg_problem_preloop": (_c_int, [_c_void_p, _c_double]),
    "b200_mg_problem_call": (_c_int, [_c_void_p, _c_double, _c_double]),
    "b200_mg_problem_postloop": (_c_int, [_c_void_p]),
    "b200_mg_problem_destroy": (_c_int, [_c_void_p]),
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

    def __init__(self, device=-1, _borrowed=None):
        self._lib = load()
        self._owned = _borrowed is None
        if _borrowed is not None:  # a team member's context (owned by the team)
            self._h = _borrowed
            return
        h = ctypes.c_void_p()
        _check(self._lib.b200_ctx_create(ctypes.byref(h), device))
        self._h = h

    def close(self):
        if getattr(self, "_h", None):
            if self._owned:
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

    def alloc_device(self, nbytes):
        d = ctypes.c_void_p()
        _check(self._lib.b200_alloc_device(self._h, nbytes, ctypes.byref(d)))
        return d.value

    def free_device(self, dev):
        _check(self._lib.b200_free_device(self._h, dev))

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

    def sgemm_prepare(self, m, n, k, lda, ldb):
        _check(self._lib.b200_sgemm_prepare(self._h, m, n, k, lda, ldb))

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


def mg_partition(total, parts, index, align=1):
    """(start, count) of part `index` -- the library's own split (pure host arithmetic)."""
    a, b = _c_ll(), _c_ll()
    _check(load().b200_mg_partition(total, parts, index, align, ctypes.byref(a), ctypes.byref(b)))
    return a.value, b.value


def mg_active_gemm(world, elem, m, n, k):
    return load().b200_mg_active_gemm(world, elem, m, n, k)


def mg_active_gemv(world, elem, m, n):
    return load().b200_mg_active_gemv(world, elem, m, n)


class Team:
    """View of a b200_mg: a LOCAL team (Team(ngpus=G): this process drives every GPU, what the
    harness does with B200_NGPUS=G) or one RANK of a team (Team(rank=r, world=W, device=d); the
    64-byte IPC handles travel through `exchange`, e.g. torch.distributed.all_gather_object)."""

    def __init__(self, ngpus=None, rank=None, world=None, device=None):
        self._lib = load()
        h = ctypes.c_void_p()
        if ngpus is not None:
            _check(self._lib.b200_mg_create(ctypes.byref(h), ngpus))
            self.world, self.rank = ngpus, None
        else:
            _check(self._lib.b200_mg_create_rank(ctypes.byref(h), device if device is not None else rank, rank, world))
            self.world, self.rank = world, rank
        self._h = h
        self._mapped = []

    def close(self):
        if getattr(self, "_h", None):
            for p in self._mapped:
                self._lib.b200_mg_unmap(self._h, p)
            self._mapped = []
            self._lib.b200_mg_destroy(self._h)
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def ctx(self, member=None):
        c = ctypes.c_void_p()
        _check(self._lib.b200_mg_ctx(self._h, self.rank if member is None else member, ctypes.byref(c)))
        return Context(_borrowed=c)

    def sync(self):
        _check(self._lib.b200_mg_sync(self._h))

    # -- rank teams: IPC
    def export(self, dev=None):
        buf = ctypes.create_string_buffer(MG_HANDLE_BYTES)
        _check(self._lib.b200_mg_export(self._h, dev, buf))
        return buf.raw

    def attach_flags(self, exchange):
        """exchange(bytes) -> list of every rank's bytes (an all-gather)."""
        handles = exchange(self.export(None))
        for r, hnd in enumerate(handles):
            if r != self.rank:
                _check(self._lib.b200_mg_import(self._h, r, hnd, None))

    def map_peers(self, dev, exchange):
        """All-gather the handle of my buffer `dev`; returns every member's address of it in
        THIS process ([rank] = dev itself)."""
        handles = exchange(self.export(dev))
        out = []
        for r, hnd in enumerate(handles):
            if r == self.rank:
                out.append(dev)
                continue
            p = ctypes.c_void_p()
            _check(self._lib.b200_mg_import(self._h, r, hnd, ctypes.byref(p)))
            self._mapped.append(p.value)
            out.append(p.value)
        return out

    def map_root(self, dev, exchange, root=0):
        """Member `root`'s buffer `dev` (None elsewhere) at an address valid in this process."""
        handles = exchange(self.export(dev) if self.rank == root else b"")
        if self.rank == root:
            return dev
        p = ctypes.c_void_p()
        _check(self._lib.b200_mg_import(self._h, root, handles[root], ctypes.byref(p)))
        self._mapped.append(p.value)
        return p.value

    # -- one member's step (rank teams) / every member's step (local teams: members=[dict, ...])
    @staticmethod
    def _gemm_args(dtype, active, m, n_slab, k, alpha, A_rep, lda, dB, ldb, beta, dC, ldc, phases,
                   hA_share=None, hB=None, hC=None, C_gather=None):
        arr = (ctypes.c_void_p * len(A_rep))(*[_ptr(a) for a in A_rep])
        a = MgGemmArgs(4 if dtype == "f32" else 8, active, m, n_slab, k, alpha, beta,
                       ctypes.cast(arr, _pp), lda, _ptr(hA_share), _ptr(hB), _ptr(dB), ldb, _ptr(hC), _ptr(dC), ldc,
                       _ptr(C_gather), phases)
        a._keep = arr
        return a

    @staticmethod
    def _gemv_args(dtype, active, m_blk, n, alpha, dA, lda, x_rep, beta, dy, phases,
                   hA=None, h_lda=0, hx_share=None, hy=None, y_gather=None):
        arr = (ctypes.c_void_p * len(x_rep))(*[_ptr(a) for a in x_rep])
        a = MgGemvArgs(4 if dtype == "f32" else 8, active, m_blk, n, alpha, beta, _ptr(dA), lda, _ptr(hA), h_lda,
                       ctypes.cast(arr, _pp), _ptr(hx_share), _ptr(hy), _ptr(dy), _ptr(y_gather), phases)
        a._keep = arr
        return a

    def gemm(self, member, *args, **kw):
        a = self._gemm_args(*args, **kw)
        _check(self._lib.b200_mg_gemm(self._h, member, ctypes.byref(a)))

    def gemv(self, member, *args, **kw):
        a = self._gemv_args(*args, **kw)
        _check(self._lib.b200_mg_gemv(self._h, member, ctypes.byref(a)))

    def gemm_team(self, members):
        """members[g] = dict of the arguments of gemm() for member g (local teams)."""
        items = [self._gemm_args(**kw) for kw in members]
        arr = (MgGemmArgs * len(items))(*items)
        _check(self._lib.b200_mg_gemm_team(self._h, arr))

    def gemv_team(self, members):
        items = [self._gemv_args(**kw) for kw in members]
        arr = (MgGemvArgs * len(items))(*items)
        _check(self._lib.b200_mg_gemv_team(self._h, arr))

    # -- whole problems (local teams): the five virtuals of the backend class
    def problem(self, kind, dtype, mode, m, n, k=0):
        return Problem(self, kind, dtype, mode, m, n, k)


class Problem:
    def __init__(self, team, kind, dtype, mode, m, n, k):
        self._lib, self._team = team._lib, team
        h = [ctypes.c_void_p() for _ in range(3)]
        p = ctypes.c_void_p()
        _check(self._lib.b200_mg_problem_create(team._h, kind, 4 if dtype == "f32" else 8, mode, m, n, k,
                                                ctypes.byref(h[0]), ctypes.byref(h[1]), ctypes.byref(h[2]),
                                                ctypes.byref(p)))
        self._p = p
        counts = (m * k, k * n, m * n) if kind == MG_GEMM else (m * n, n, m)
        self.host = [host_view(x.value, c, dtype) if c else None for x, c in zip(h, counts)]

    @property
    def active(self):
        return self._lib.b200_mg_problem_active(self._p)

    def preloop(self, beta=0.0):
        _check(self._lib.b200_mg_problem_preloop(self._p, beta))

    def call(self, alpha=1.0, beta=0.0):
        _check(self._lib.b200_mg_problem_call(self._p, alpha, beta))

    def postloop(self):
        _check(self._lib.b200_mg_problem_postloop(self._p))

    def destroy(self):
        if self._p:
            _check(self._lib.b200_mg_problem_destroy(self._p))
            self._p = None


def host_view(ptr, count, dtype):
    """numpy view of `count` elements of pinned/managed host memory at `ptr`."""
    import numpy as np
    ct = {"f32": ctypes.c_float, "f64": ctypes.c_double}[dtype]
    buf = (ct * count).from_address(ptr)
    return np.frombuffer(buf, dtype=np.float32 if dtype == "f32" else np.float64, count=count)
