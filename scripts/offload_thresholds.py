#!/usr/bin/env python
"""Offload thresholds from harness CSVs, computed by the reference's UNMODIFIED
calculateOffloadThreshold.py (run through the texttable stand-in) and summarised
side by side for the cuBLAS and B200 backends.

    python scripts/offload_thresholds.py gpurun_out/<sweep>/cublas_dense gpurun_out/<sweep>/b200_dense

Each harness CSV holds cpu + 3 GPU rows per size; the reference script wants a
CPU-only and a GPU-only file (reference calculateOffloadThreshold.py:71-77).
"""
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GPU_BLOB_REFERENCE", "/root/reference")
SHIMS = os.path.join(ROOT, "gpu-blas-offload-benchmark_b200", "harness", "pyshims")


def split(csv_path, tmp):
    lines = open(csv_path).read().splitlines()
    hdr, rows = lines[0], lines[1:]
    cpu = [r for r in rows if r.startswith("cpu,")]
    gpu = [r for r in rows if r.startswith("gpu_")]
    c, g = os.path.join(tmp, "cpu.csv"), os.path.join(tmp, "gpu.csv")
    open(c, "w").write("\n".join([hdr] + cpu) + "\n")
    open(g, "w").write("\n".join([hdr] + gpu) + "\n")
    return c, g


def thresholds(csv_path):
    """-> {mode: 'M,N,K' or 'N/A'} parsed from the reference script's table."""
    with tempfile.TemporaryDirectory() as tmp:
        c, g = split(csv_path, tmp)
        env = dict(os.environ, PYTHONPATH=SHIMS)
        r = subprocess.run([sys.executable, os.path.join(REF, "calculateOffloadThreshold.py"), c, g],
                           capture_output=True, text=True, env=env)
    out = {}
    for line in r.stdout.splitlines():
        cells = [x.strip() for x in line.strip().strip("|").split("|")]
        if len(cells) >= 4 and cells[0] in ("Once", "Always", "Unified", "Offload Once", "Offload Always"):
            out[cells[0]] = cells[1:]
    return out, r.stdout, r.stderr


def main():
    dirs = sys.argv[1:]
    names = sorted(f for f in os.listdir(dirs[0]) if f.endswith(".csv"))
    for f in names:
        print("==", f)
        for d in dirs:
            p = os.path.join(d, f)
            if not os.path.exists(p):
                continue
            _, txt, err = thresholds(p)
            label = os.path.basename(d.rstrip("/"))
            body = txt.strip() or err.strip()[-300:]
            print(f"[{label}]")
            print(body)


if __name__ == "__main__":
    main()
