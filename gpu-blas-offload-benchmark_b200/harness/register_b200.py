"""Register the B200 backend in a scratch copy of the GPU-BLOB tree.

GPU-BLOB selects its backends at compile time through #if/#elif include chains
(include/doGemm.hh:21-27, include/doGemv.hh:21-27, include/utilities.hh:25-35)
and a Makefile branch per library.  A new backend is therefore registered by
three one-line #elif insertions plus its own directory next to cuBLAS/ -- this
script applies exactly those insertions to a COPY of the reference tree (the
reference checkout is read-only and nothing from it is stored in this repo).
INTEGRATION.md shows the same change as a patch for a maintainer.
"""
import os
import shutil

GEMM_ANCHOR = '#elif defined GPU_ROCBLAS\n#include "../rocBLAS/gemm.hh"\n'
GEMV_ANCHOR = '#elif defined GPU_ROCBLAS\n#include "../rocBLAS/gemv.hh"\n'
NAME_ANCHOR = '#elif defined GPU_ROCBLAS\n#define GPU_LIB_NAME "AMD rocBLAS"\n'


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
    return dest
