"""Register the B200 backend in a scratch copy of the GPU-BLOB tree.

GPU-BLOB selects its backends at compile time through #if/#elif include chains
(include/doGemm.hh:21-27, include/doGemv.hh:21-27, include/utilities.hh:25-35)
and a Makefile branch per library (Makefile:167-217; unknown GPU_LIB values are
rejected at :215-216, -DGPU_$(GPU_LIB) comes from :224-226).  A new backend is
therefore registered by three one-line #elif insertions, one Makefile branch
and its own directory next to cuBLAS/ -- this script applies exactly those
insertions to a COPY of the reference tree (the reference checkout is read-only
and nothing from it is stored in this repo), and build.py then builds the copy
with the reference's own `make COMPILER=GNU CPU_LIB=OPENBLAS GPU_LIB=B200`.
INTEGRATION.md shows the same change as a patch for a maintainer.
"""
import os
import shutil

GEMM_ANCHOR = '#elif defined GPU_ROCBLAS\n#include "../rocBLAS/gemm.hh"\n'
GEMV_ANCHOR = '#elif defined GPU_ROCBLAS\n#include "../rocBLAS/gemv.hh"\n'
NAME_ANCHOR = '#elif defined GPU_ROCBLAS\n#define GPU_LIB_NAME "AMD rocBLAS"\n'
# the closing branch of the GPU_LIB chain: the B200 branch goes right before it
MAKE_ANCHOR = 'else\n$(warning Provided GPU_LIB not valid (use CUBLAS, ONEMKL, ROCBLAS). No GPU kernels will be run.)\n'
MAKE_BRANCH = '''else ifeq ($(GPU_LIB), B200)
# B200-native backend: the host side is plain C++ over libb200blas.so (no CUDA headers, any host compiler)
$(warning Users may be required to do the following to use $(COMPILER) with $(GPU_LIB):)
$(info $(TAB)$(TAB)Add `CXXFLAGS="-I<B200BLAS_DIR>/include -L<B200BLAS_DIR>/lib -Wl,-rpath,<B200BLAS_DIR>/lib"` to make command)
$(info )
override CXXFLAGS += -lb200blas
HEADER_FILES += $(wildcard B200/*.hh)

'''


def _insert_before(path, anchor, addition):
    with open(path) as f:
        text = f.read()
    if addition in text:
        return
    if anchor not in text:
        raise RuntimeError(f"{path}: registration anchor not found -- reference layout changed")
    with open(path, "w") as f:
        f.write(text.replace(anchor, addition + anchor, 1))


def _insert_after(path, anchor, addition):
    with open(path) as f:
        text = f.read()
    if addition in text:
        return
    if anchor not in text:
        raise RuntimeError(f"{path}: registration anchor not found -- reference layout changed")
    with open(path, "w") as f:
        f.write(text.replace(anchor, anchor + addition, 1))


def make_overlay(reference_root, dest, b200_dir):
    """Copy `reference_root` to `dest`, add B200/ and the three registrations."""
    shutil.copytree(reference_root, dest, ignore=shutil.ignore_patterns(".git"))
    for root, dirs, files in os.walk(dest):
        for d in dirs:
            os.chmod(os.path.join(root, d), 0o755)
        for fn in files:
            os.chmod(os.path.join(root, fn), 0o644)
    shutil.copytree(b200_dir, os.path.join(dest, "B200"))
    inc = os.path.join(dest, "include")
    _insert_after(os.path.join(inc, "doGemm.hh"), GEMM_ANCHOR,
                  '#elif defined GPU_B200\n#include "../B200/gemm.hh"\n')
    _insert_after(os.path.join(inc, "doGemv.hh"), GEMV_ANCHOR,
                  '#elif defined GPU_B200\n#include "../B200/gemv.hh"\n')
    _insert_after(os.path.join(inc, "utilities.hh"), NAME_ANCHOR,
                  '#elif defined GPU_B200\n#define GPU_LIB_NAME "B200 native (libb200blas)"\n')
    _insert_before(os.path.join(dest, "Makefile"), MAKE_ANCHOR, MAKE_BRANCH)
    return dest
