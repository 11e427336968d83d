"""Build recipe for libb200blas.so (hand-written CUDA for sm_100a) and for the
GPU-BLOB harness binary that links the B200 backend.

Everything is built IN-TREE so the artefacts travel to the GPU box with the
repo snapshot:

    gpu-blas-offload-benchmark_b200/lib/libb200blas.so
    gpu-blas-offload-benchmark_b200/lib/b200_parity         (C++ backend parity driver)
    gpu-blas-offload-benchmark_b200/lib/gpu-blob-b200       (reference harness + B200 backend)
    gpu-blas-offload-benchmark_b200/lib/libconsume.so       (the harness' optimisation barrier)

    python gpu-blas-offload-benchmark_b200/build.py [--force] [--no-harness]
"""
import hashlib
import os
import shutil
import subprocess
import sys
import tempfile

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
LIB = os.path.join(LIBDIR, "libb200blas.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
REFERENCE = os.environ.get("GPU_BLOB_REFERENCE", "/root/reference")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint(paths):
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd, **kw):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, **kw)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("command failed: " + " ".join(cmd))
    return r.stdout


def build_library(force=False, verbose=False):
    """nvcc -> lib/libb200blas.so; per-file objects cached on a content hash."""
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "b200blas.h"))
    objs, rebuilt = [], False
    log = []
    for src in _sources():
        base = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(objdir, base + ".o")
        stamp = obj + ".sha"
        fp = _fingerprint([src] + headers)
        if force or not os.path.exists(obj) or not os.path.exists(stamp) or open(stamp).read() != fp:
            out = _run([NVCC] + NVCC_FLAGS + ["-c", src, "-o", obj])
            log.append(out)
            with open(stamp, "w") as f:
                f.write(fp)
            rebuilt = True
        objs.append(obj)
    if rebuilt or not os.path.exists(LIB):
        _run([NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a",
                                                     "-cudart", "static", "-ldl"])
    if log:
        with open(os.path.join(LIBDIR, "ptxas.log"), "w") as f:
            f.write("\n".join(log))
        if verbose:
            print("\n".join(log))
    return LIB


def _load_register():
    import importlib.util
    spec = importlib.util.spec_from_file_location("register_b200", os.path.join(PKG, "harness", "register_b200.py"))
    reg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(reg)
    return reg


def build_harness(force=False):
    """The reference harness with the B200 backend registered (lib/gpu-blob-b200)
    and the C++ parity driver over the same backend classes (lib/b200_parity).

    /root/reference is read-only and its quote-includes resolve relative to the
    including file, so registration is a build-time overlay: copy the tree to a
    scratch dir, apply harness/register_b200.py (three #elif insertions -- the
    same places cuBLAS/rocBLAS/oneMKL are registered), drop B200/ next to
    cuBLAS/, compile.  Only the binaries come back; without /root/reference (GPU
    box) the prebuilt ones are used.
    """
    blob = os.path.join(LIBDIR, "gpu-blob-b200")
    parity = os.path.join(LIBDIR, "b200_parity")
    if not os.path.isdir(os.path.join(REFERENCE, "include")):
        return [p if os.path.exists(p) else None for p in (blob, parity)]
    consume = os.path.join(ROOT, "oracle", "_ref", "libconsume.so")
    if not os.path.exists(consume):
        _run(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    deps = [os.path.join(PKG, "B200", f) for f in os.listdir(os.path.join(PKG, "B200"))]
    deps += [os.path.join(PKG, "harness", "register_b200.py"),
             os.path.join(PKG, "harness", "parity_driver.cc"),
             os.path.join(ROOT, "include", "b200blas.h")]
    fresh = all(os.path.exists(o) and all(os.path.getmtime(o) >= os.path.getmtime(d) for d in deps)
                for o in (blob, parity))
    if fresh and not force:
        return [blob, parity]
    reg = _load_register()
    scratch = tempfile.mkdtemp(prefix="gpu-blob-b200-")
    try:
        tree = os.path.join(scratch, "tree")
        reg.make_overlay(REFERENCE, tree, os.path.join(PKG, "B200"))
        sp = subprocess.check_output(
            [sys.executable, "-c", "import scipy,os;print(os.path.dirname(os.path.dirname(scipy.__file__)))"],
            text=True).strip()
        scipy_libs = os.path.join(sp, "scipy.libs")
        openblas = [os.path.join(scipy_libs, f) for f in os.listdir(scipy_libs) if f.startswith("libscipy_openblas")][0]
        # ---- the harness: the reference's own build entry point on the overlay,
        #   make COMPILER=GNU CPU_LIB=OPENBLAS GPU_LIB=B200 CXXFLAGS="-I.. -L.. -Wl,-rpath,.."
        # (reference Makefile:44-60,155-159,167-226).  Two command-line overrides, both about WHERE
        # the binary runs, not what it is: CXXFLAGS_GNU swaps -march=native for x86-64-v3 (built in a
        # CPU container, run on the GPU box), and the rpath is $ORIGIN so lib/ is self-contained.
        shim = os.path.join(scratch, "shim")
        os.makedirs(shim)
        os.symlink(openblas, os.path.join(shim, "libopenblas.so"))
        gfortran = [os.path.join(scipy_libs, f) for f in os.listdir(scipy_libs) if f.startswith("libgfortran")][0]
        os.symlink(gfortran, os.path.join(shim, "libgfortran.so"))
        cxxflags = " ".join(["-I" + os.path.join(ROOT, "oracle", "cblas_shim"), "-L" + shim,
                             "-Wl,-rpath," + scipy_libs, "-I" + os.path.join(ROOT, "include"), "-L" + LIBDIR,
                             "-Wl,-rpath,\\$$ORIGIN"])
        out = _run(["make", "-C", tree, "COMPILER=GNU", "CPU_LIB=OPENBLAS", "GPU_LIB=B200",
                    "CXXFLAGS_GNU=-std=c++17 -Wall -Ofast -march=x86-64-v3", "CXXFLAGS=" + cxxflags])
        with open(os.path.join(LIBDIR, "harness_make.log"), "w") as f:
            f.write(out)
        shutil.copy(os.path.join(tree, "gpu-blob"), blob)
        shutil.copy(os.path.join(tree, "src", "Consume", "libconsume.so"), os.path.join(LIBDIR, "libconsume.so"))
        common = ["-std=c++17", "-Wall", "-march=x86-64-v3", "-DGPU_B200", "-I" + os.path.join(ROOT, "include")]
        link = ["-L" + LIBDIR, "-lb200blas", "-Wl,-rpath,$ORIGIN"]
        _run(["g++", os.path.join(PKG, "harness", "parity_driver.cc"), "-O2", "-I" + tree] + common + link +
             ["-o", parity])
    finally:
        shutil.rmtree(scratch, ignore_errors=True)
    return [blob, parity]


def build_all(force=False, harness=True, verbose=False):
    lib = build_library(force=force, verbose=verbose)
    extras = build_harness(force=force) if harness else []
    return lib, extras


if __name__ == "__main__":
    lib, extras = build_all(force="--force" in sys.argv, harness="--no-harness" not in sys.argv,
                            verbose="-v" in sys.argv)
    print("built", lib)
    for e in extras:
        print("built" if e else "skipped", e)
