#!/usr/bin/env python
"""Offload-threshold comparison of two harness sweeps (reference cuBLAS backend vs B200 backend)
from their CSVs, with the reference's own rule (include/doGemm.hh:418-490: a threshold is set at
the first size where GPU GFLOP/s > CPU GFLOP/s and reset when the CPU wins two sizes in a row).

Two columns per backend:
  in-run      CPU series of that backend's own harness run (what the harness prints)
  common CPU  both backends judged against ONE CPU series: per size, the faster of the two runs'
              CPU measurements (same box, same OpenBLAS binary; the CPU rows of a 16-thread
              OpenBLAS at these sizes are noisy enough to move a threshold by +-50)

usage: thresholds_md.py <sweep dir with cublas_dense/ b200_dense/ [cublas_d*/ b200_d*/]>
"""
import csv
import glob
import os
import sys

MODES = [("gpu_offloadOnce", "Offload Once"), ("gpu_offloadAlways", "Offload Always"), ("gpu_unified", "Unified Memory")]


def load(path):
    rows = {}
    order = []
    for r in csv.DictReader(open(path)):
        key = (int(r["M"]), int(r["N"]), int(r["K"]))
        if key not in rows:
            rows[key] = {}
            order.append(key)
        rows[key][r["Device"]] = float(r["GFLOP/s"])
    return order, rows


def threshold(order, cpu, gpu):
    thr, prev = None, 0.0
    for key in order:
        c, g = cpu[key], gpu[key]
        if thr is not None and c >= g and c >= prev:
            thr = None
        if thr is None and c < g:
            thr = key
        prev = g
    return thr


def fmt(t, gemv):
    if t is None:
        return "-"
    return f"{t[0]}x{t[1]}" if gemv else f"{t[0]}x{t[1]}x{t[2]}"


def main():
    d = sys.argv[1]
    names = sorted(os.path.basename(p) for p in glob.glob(os.path.join(d, "b200_dense", "*.csv")))
    print("| kernel | problem type | mode | cuBLAS backend (in-run) | B200 backend (in-run) | cuBLAS (common CPU) | B200 (common CPU) |")
    print("|---|---|---|---|---|---|---|")
    wins = ties = losses = 0
    for nm in names:
        pc, pb = os.path.join(d, "cublas_dense", nm), os.path.join(d, "b200_dense", nm)
        if not os.path.exists(pc):
            continue
        oc, rc = load(pc)
        ob, rb = load(pb)
        gemv = "gemv" in nm
        common = {k: max(rc[k]["cpu"], rb[k]["cpu"]) for k in oc if k in rb}
        order = [k for k in oc if k in rb]
        kern, ptype = nm.split("_", 1)
        for dev, label in MODES:
            t_c = threshold(oc, {k: rc[k]["cpu"] for k in oc}, {k: rc[k][dev] for k in oc})
            t_b = threshold(ob, {k: rb[k]["cpu"] for k in ob}, {k: rb[k][dev] for k in ob})
            u_c = threshold(order, common, {k: rc[k][dev] for k in order})
            u_b = threshold(order, common, {k: rb[k][dev] for k in order})
            if t_c is None and t_b is None and u_c is None and u_b is None:
                continue
            print(f"| {kern.upper()} | {ptype[:-4]} | {label} | {fmt(t_c, gemv)} | {fmt(t_b, gemv)} | {fmt(u_c, gemv)} | {fmt(u_b, gemv)} |")
            size = lambda t: float("inf") if t is None else t[0] * t[1] * (t[2] if not gemv else 1)
            if size(u_b) < size(u_c):
                wins += 1
            elif size(u_b) == size(u_c):
                ties += 1
            else:
                losses += 1
    print(f"\nCommon-CPU comparison: B200 threshold lower in {wins} rows, equal in {ties}, higher in {losses}.")
    pts = sorted(int(os.path.basename(p)[6:]) for p in glob.glob(os.path.join(d, "b200_d[0-9]*")) if os.path.isdir(p))
    if pts:
        print("\n## Sampled large points (`-s d -d d`, 10 iterations, harness wall clock incl. transfers), GFLOP/s\n")
        print("| d | file | shape | cpu | once cuBLAS | once B200 | always cuBLAS | always B200 | unified cuBLAS | unified B200 |")
        print("|---|---|---|---|---|---|---|---|---|---|")
        for p in pts:
            for nm in sorted(os.listdir(os.path.join(d, f"b200_d{p}"))):
                pc, pb = os.path.join(d, f"cublas_d{p}", nm), os.path.join(d, f"b200_d{p}", nm)
                if not os.path.exists(pc):
                    continue
                oc, rc = load(pc)
                ob, rb = load(pb)
                for k in ob:
                    if k not in rc:
                        continue
                    print(f"| {p} | {nm[:-4]} | {'x'.join(map(str, k if 'gemm' in nm else k[:2]))} | {max(rc[k]['cpu'], rb[k]['cpu']):.0f} | "
                          + " | ".join(f"{rc[k][dev]:.0f} | {rb[k][dev]:.0f}" for dev, _ in MODES) + " |")


if __name__ == "__main__":
    main()
