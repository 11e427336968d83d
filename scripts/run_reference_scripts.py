#!/usr/bin/env python
"""Run the reference's two post-processing scripts UNCHANGED on a harness CSV directory:

    python scripts/run_reference_scripts.py <CSV dir> [--out <dir for Graphs_* and tables>]

calculateOffloadThreshold.py (reference :1-120) wants a CPU-only and a GPU-only CSV per kernel
and `texttable`; createGflopsGraphs.py (reference :1-380) wants `matplotlib`.  Neither module is
installed here, so harness/pyshims (appended BEHIND site-packages) supplies stand-ins; with the
real modules installed they are never imported.  Needs /root/reference (build container only).
"""
import glob
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("GPU_BLOB_REFERENCE", "/root/reference")
SHIMS = os.path.join(ROOT, "gpu-blas-offload-benchmark_b200", "harness", "pyshims")


def main():
    csv_dir = os.path.abspath(sys.argv[1])
    out = os.path.abspath(sys.argv[sys.argv.index("--out") + 1]) if "--out" in sys.argv else os.getcwd()
    os.makedirs(out, exist_ok=True)
    env = dict(os.environ)
    # behind site-packages: python puts PYTHONPATH before site-packages, so only add the shims
    # for modules that are really missing
    need = [m for m in ("texttable", "matplotlib") if subprocess.call(
        [sys.executable, "-c", f"import {m}"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)]
    if need:
        env["PYTHONPATH"] = SHIMS + os.pathsep + env.get("PYTHONPATH", "")
        print(f"[run_reference_scripts] stand-ins in use for: {', '.join(need)}")
    rel = os.path.relpath(csv_dir, out)
    rc = subprocess.call([sys.executable, os.path.join(REF, "createGflopsGraphs.py"), rel], cwd=out, env=env)
    print(f"[run_reference_scripts] createGflopsGraphs.py rc={rc}")
    with tempfile.TemporaryDirectory() as tmp, open(os.path.join(out, "offload_thresholds.txt"), "w") as log:
        for path in sorted(glob.glob(os.path.join(csv_dir, "*.csv"))):
            lines = open(path).read().splitlines()
            hdr, rows = lines[0], lines[1:]
            c, g = os.path.join(tmp, "cpu.csv"), os.path.join(tmp, "gpu.csv")
            open(c, "w").write("\n".join([hdr] + [r for r in rows if r.startswith("cpu,")]) + "\n")
            open(g, "w").write("\n".join([hdr] + [r for r in rows if r.startswith("gpu_")]) + "\n")
            r = subprocess.run([sys.executable, os.path.join(REF, "calculateOffloadThreshold.py"), c, g],
                               env=env, capture_output=True, text=True)
            log.write(f"== {os.path.basename(path)} (rc={r.returncode})\n{r.stdout}{r.stderr}\n")
            rc |= r.returncode
    print(f"[run_reference_scripts] calculateOffloadThreshold.py over {len(glob.glob(os.path.join(csv_dir, '*.csv')))} files -> "
          f"{os.path.join(out, 'offload_thresholds.txt')}, overall rc={rc}")
    return rc


if __name__ == "__main__":
    sys.exit(main())
