"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports
every symbol include/b200blas.h declares, the ctypes table matches the header,
argument errors are reported (not crashed on), and -- with no GPU -- the
product refuses to run instead of falling back to a CPU path."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200blas.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ("b200_ctx_create", "b200_alloc", "b200_h2d_async", "b200_d2h_async",
                 "b200_prefetch_async", "b200_sync", "b200_sgemm", "b200_dgemm", "b200_sgemv",
                 "b200_dgemv", "b200_dgemm_offload", "b200_last_error"):
        assert must in syms


def test_library_exports_every_declared_symbol(blob):
    lib = ctypes.CDLL(blob.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_table_matches_header(blob):
    assert sorted(blob.SIGNATURES) == declared_symbols()


def test_offload_enum_matches_reference(blob):
    # include/utilities.hh:50-54: always = 0, once, unified
    assert (blob.OFFLOAD_ALWAYS, blob.OFFLOAD_ONCE, blob.OFFLOAD_UNIFIED) == (0, 1, 2)
    hdr = open(os.path.join(ROOT, "include", "b200blas.h")).read()
    assert "B200_OFFLOAD_ALWAYS = 0, B200_OFFLOAD_ONCE = 1, B200_OFFLOAD_UNIFIED = 2" in hdr


def test_sass_is_sm100a_native(blob):
    """The shipped cubin is sm_100a and contains the FP64 tensor instruction."""
    out = subprocess.run(["cuobjdump", "-lelf", blob.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    sass = subprocess.run(["cuobjdump", "-sass", blob.LIB_PATH], capture_output=True, text=True).stdout
    assert "DMMA" in sass
    # the Blackwell machinery the kernels are written for, by SASS mnemonic (profiles/r2_sass_mnemonics.txt):
    # tcgen05.mma / tcgen05.ld / tcgen05.commit, TMA load and TMA store, mbarrier, cluster barrier
    # (CTA pairs, tall-K clusters), setmaxnreg, elect.sync
    for mnemonic in ("UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "UCGABAR_ARV", "USETMAXREG", "ELECT"):
        assert mnemonic in sass, mnemonic
    # ... and no library GEMM: the product links cudart only
    ldd = subprocess.run(["ldd", blob.LIB_PATH], capture_output=True, text=True).stdout
    assert "cublas" not in ldd.lower() and "cutlass" not in ldd.lower(), ldd


def test_no_gpu_means_loud_failure_not_cpu_fallback(blob):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(blob.B200Error, match="no CUDA device"):
        blob.Context()


def test_product_does_not_touch_the_oracle():
    """Nothing under the product package may reference oracle/ (parity rule)."""
    pkg = os.path.join(ROOT, "gpu-blas-offload-benchmark_b200")
    bad = []
    for root, _, files in os.walk(pkg):
        if os.sep + "lib" in root:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".hh", ".h", ".cc")):
                text = open(os.path.join(root, fn)).read()
                if re.search(r"bloboracle|blob_oracle|oracle_lib|blobref_", text):
                    bad.append(os.path.join(root, fn))
    assert not bad, bad
