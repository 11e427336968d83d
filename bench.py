#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 backend for GPU-BLOB.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload dgemm|sgemm|dgemv|sgemv] [--dim D] [--no-extras] [--weak]

One "step" is one pass of the hot path over one batch of synthetic input: one
GEMM (or GEMV) call of the named workload.  Metric = GFLOP/s with the
reference's own FLOP formulas (include/doGemm.hh:493-505 -> 2MNK + MN;
include/doGemv.hh:377-388 -> 2MN + M), as BASELINE.json asks.

N = 1   square DGEMM M=N=K=8192, the top of BASELINE.json configs[1]
        ("square DGEMM/SGEMM M=N=K 1->8192 on 1 B200").
N > 1   (torchrun, one rank per GPU) BASELINE.json configs[4]: ONE square problem
        M=N=K=32768 partitioned over the N GPUs ("scaling": "strong") through the library's
        multi-GPU team (csrc/mg.cu, one rank per process, peers mapped by CUDA IPC): rank r
        owns a column slab of B and C and a K-share of A; every step the shares are pushed to
        all replicas of A over NVLink chunk by chunk while the GEMM consumes them, and the
        C slabs are gathered on rank 0 -- the exchange is INSIDE the timed step.
        --weak restores round 1's independent 8192-wide slabs (no exchange in the step).

value      device-resident: operands already in HBM (N > 1: each rank's K-share of A and its
           slab of B), CUDA events on the library's compute stream around exactly K steps,
           max over ranks.
e2e        the same step through the reference-facing plugin path with HOST buffers, copies
           inside the timed region: N = 1 b200_?gemm_offload (gpu::gemm_gpu<T>::callGemm in
           Always mode); N > 1 the team's upload + replicate + compute + download step (what
           the backend classes run with B200_NGPUS=N), every rank uploading only its share of
           A and its slab of B over its own PCIe link and downloading its slab of C.
roofline   achieved FLOP/s (or GB/s) of the dominant kernel vs the measured peak of the pipe
           that bounds it (+ the spec-derived figure).
cpu_baseline / --impl reference
           the reference's own CPU path -- cpu::gemm_cpu<T>::compute() of its OpenBLAS backend
           (include/kernels/gemm.hh:19-42, OpenBLAS/gemm.hh:25-43), compiled from
           /root/reference into oracle/_ref/libblobref.so -- on this box's host cores, at the
           size stated in config.workload.
gpu_baseline
           the reference's cuBLAS backend classes (cuBLAS/gemm.hh) through
           oracle/_ref/blobref-cublas, run as a separate process outside every timed region.
"""
import argparse
import ctypes
import importlib
import json
import os
import subprocess
import sys
import threading
import time

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")  # the team's streams never share a hardware queue

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ELEM = {"f64": 8, "f32": 4}
TOL = {"f64": 1e-12, "f32": 1e-5}
STRONG_DIM = 32768          # BASELINE.json configs[4]
CPU_REF_GEMM_DIM = 8192     # largest GEMM the reference arm runs whole (32768^3 is ~47 s per step on 16 cores)


def gemm_flops(m, n, k):  # include/doGemm.hh:493-505, BETA == 0
    return 2 * m * n * k + m * n


def gemv_flops(m, n):  # include/doGemv.hh:377-388, BETA == 0
    return 2 * m * n + m


def gemm_bytes(m, n, k, s):  # include/doGemm.hh:508-512 (A, B read; C written, beta = 0)
    return (m * k + k * n + m * n) * s


def gemv_bytes(m, n, s):  # include/doGemv.hh:391-395
    return (m * n + n + m) * s


def ncu_traffic(workload, kernel_name):
    """DRAM bytes (read + write) of ONE launch of the dominant kernel, from the committed
    `ncu --set full` capture of this workload's bench command (profiles/r2_traffic.json, else
    r1_traffic.json; written by scripts/ncu_traffic.py from the .ncu-rep); None when the capture
    is of a different kernel than the one that just ran."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(p):
            continue
        t = json.load(open(p)).get(workload)
        if t and t.get("kernel_label") == kernel_name:
            return t["dram_read_bytes"] + t["dram_write_bytes"], t.get("source")
    return None, None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "MEASURED_PEAKS.json"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


def prefix(dtype):
    return "d" if dtype == "f64" else "s"


# ---------------------------------------------------------------- reference arm

def load_reference():
    so = os.path.join(ROOT, "oracle", "_ref", "libblobref.so")
    if not os.path.exists(so):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"],
                        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    lib.blobref_blas_config.restype = ctypes.c_char_p
    i, vp = ctypes.c_int, ctypes.c_void_p
    pd = ctypes.POINTER(ctypes.c_double)
    for name in ("blobref_dgemm", "blobref_sgemm"):
        getattr(lib, name).argtypes = [i, i, i, i, vp, pd, pd]
    for name in ("blobref_dgemv", "blobref_sgemv"):
        getattr(lib, name).argtypes = [i, i, i, vp, pd, pd]
    return lib


def reference_dims(kind, dim):
    """The size the CPU arm runs WHOLE.  GEMM is capped at 8192^3 (config 5's 32768^3 is ~47 s
    per step on the box's cores); the cap is stated in config.workload, never hidden."""
    return min(dim, CPU_REF_GEMM_DIM) if kind == "gemm" else dim


def time_reference(kind, dtype, dim, steps, warmup):
    """`steps` iterations of the reference's own cpu::gem?_cpu<T>::compute() -- its generator
    (srand(19123005), rand() % 100 / 7.0 and / 3.0), its cblas call + consume(), its
    wall-clock timer (include/kernels/gemm.hh:19-42) -- at M=N(=K)=dim."""
    lib = load_reference()
    cores = os.cpu_count() or 1
    lib.blobref_blas_set_threads(min(cores, 64))  # OpenBLAS build: MAX_THREADS=64
    cs, sec = ctypes.c_double(), ctypes.c_double()

    def run(iters):
        if kind == "gemm":
            fn = lib.blobref_dgemm if dtype == "f64" else lib.blobref_sgemm
            fn(dim, dim, dim, iters, None, ctypes.byref(cs), ctypes.byref(sec))
        else:
            fn = lib.blobref_dgemv if dtype == "f64" else lib.blobref_sgemv
            fn(dim, dim, iters, None, ctypes.byref(cs), ctypes.byref(sec))
        return sec.value

    if warmup > 0:
        run(warmup)
    dt = run(steps)
    flops = gemm_flops(dim, dim, dim) if kind == "gemm" else gemv_flops(dim, dim)
    return {"value": flops * steps / dt * 1e-9, "ms_per_step": dt / steps * 1e3,
            "cores": lib.blobref_blas_threads(), "kind": "reference", "checksum": cs.value,
            "sample": f"{prefix(dtype)}{kind} M=N{'=K' if kind == 'gemm' else ''}={dim}: {steps} iterations inside one "
                      f"cpu::gem{'m' if kind == 'gemm' else 'v'}_cpu<T>::compute() of the reference (its generator, "
                      f"cblas call + consume(), its timer); {lib.blobref_blas_config().decode().strip()}"}


def run_reference_arm(args, kind, dtype, dim, strong):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rdim = reference_dims(kind, dim)
    r = time_reference(kind, dtype, rdim, args.steps, args.warmup)
    cfg = workload_config(kind, dtype, dim, args.gpus, strong)
    if rdim != dim:
        cfg["workload"] = (f"{prefix(dtype)}gemm_square M=N=K={rdim} -- BOUNDED SAMPLE of the {args.gpus}-GPU workload "
                           f"M=N=K={dim} (one {dim}^3 step is ~{gemm_flops(dim, dim, dim) / (r['value'] * 1e9):.0f} s "
                           "on these host cores); the rate, not the size, is what carries over")
        cfg["sampled"] = True
    elif args.gpus > 1:
        cfg["workload"] = f"{prefix(dtype)}{kind}_square M=N{'=K' if kind == 'gemm' else ''}={rdim} (the whole problem the {args.gpus} GPUs partition)"
    line = {
        "impl": "reference",
        "metric": f"{prefix(dtype).upper()}{kind.upper()} GFLOP/s (GPU-BLOB formula)",
        "value": r["value"], "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if (strong and args.gpus > 1) else "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"],
                         "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "checksum": r["checksum"],
    }
    print(json.dumps(line), flush=True)


def cublas_baseline(kind, dtype, dim):
    """The reference's cuBLAS backend (its own classes, its own timer) in a separate process:
    Once mode = kernel rate with resident operands, Always mode = its end-to-end rate."""
    exe = os.path.join(ROOT, "oracle", "_ref", "blobref-cublas")
    if not os.path.exists(exe):
        return {"unavailable": "oracle/_ref/blobref-cublas not built"}
    out = {"impl": "reference cuBLAS backend (cuBLAS/gemm.hh, cuBLAS/gemv.hh) via oracle/_ref/blobref-cublas",
           "unit": "GFLOP/s", "size": dim,
           "note": "its SGEMM runs under CUBLAS_TENSOR_OP_MATH (cuBLAS/gemm.hh:53)"}
    dims = [str(dim)] * (3 if kind == "gemm" else 2)
    for mode, iters in (("once", 10), ("always", 5)):
        try:
            r = subprocess.run([exe, prefix(dtype) + kind, mode] + dims + [str(iters), "2"], capture_output=True,
                               text=True, timeout=300, cwd=os.path.dirname(exe))
            rows = [ln.split() for ln in r.stdout.strip().splitlines() if len(ln.split()) == 3]
            if r.returncode != 0 or not rows:
                out[mode] = {"error": (r.stdout + r.stderr)[-200:]}
                continue
            # second repeat: the handle, the streams and the allocator are warm (the harness keeps one object too)
            out[mode] = {"gflops": float(rows[-1][2]), "seconds": float(rows[-1][1]), "iterations": iters,
                         "checksum": float(rows[-1][0])}
        except (OSError, subprocess.TimeoutExpired) as ex:
            out[mode] = {"error": str(ex)[-200:]}
    return out


# ------------------------------------------------------------------- own arm

def workload_config(kind, dtype, dim, gpus, strong=True):
    p = prefix(dtype)
    if gpus > 1 and strong:
        wl = (f"{p}gemm M=N=K={dim} in {gpus} column slabs (~{dim // gpus} columns of B and C and a K-share of "
              f"~{dim // gpus} columns of A per GPU; A replicated over NVLink and C gathered inside every step)"
              if kind == "gemm" else
              f"{p}gemv M=N={dim} in {gpus} row blocks (~{dim // gpus} rows of A and y per GPU; x replicated over "
              "NVLink and y gathered inside every step)")
        base = "configs[4]"
    elif kind == "gemm":
        wl = f"{p}gemm_square M=N=K={dim}" if gpus == 1 else \
            f"{p}gemm column slabs: M=K={dim}, N={dim}x{gpus} ({dim} columns of B and C per GPU, A replicated)"
        base = "configs[1]"
    else:
        wl = f"{p}gemv_square M=N={dim}" if gpus == 1 else \
            f"{p}gemv row blocks: N={dim}, M={dim}x{gpus} ({dim} rows of A and y per GPU, x replicated)"
        base = "configs[3]"
    return {"workload": wl, "baseline_config": base,
            "alpha": 1, "beta": 0, "layout": "column-major, NoTrans",
            "l2": "operands larger than the 126 MB L2 (no flush needed)",
            "inputs": "rand()%100 / 7.0 and / 3.0 value distribution of the reference generator"}


def device_inputs(torch, kind, dtype, dim, seed, loc=None):
    """loc = this rank's extent of the partitioned dimension (weak scaling / one GPU: dim)."""
    loc = loc or dim
    tdt = torch.float64 if dtype == "f64" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(seed)

    def fill(count, div):
        return torch.randint(0, 100, (count,), generator=g, device="cuda").to(tdt) / div
    if kind == "gemm":
        return fill(dim * dim, 7.0), fill(dim * loc, 3.0), torch.zeros(dim * loc, dtype=tdt, device="cuda")
    return fill(loc * dim, 7.0), fill(dim, 3.0), torch.zeros(loc, dtype=tdt, device="cuda")


def host_fill(view, div, seed):
    """Reference value distribution into a (large) pinned buffer: one random block of a prime
    length, repeated -- fast, and a misplaced chunk would not line up with the repetition."""
    import numpy as np
    period = min(view.size, (1 << 23) + 11)
    block = (np.random.default_rng(seed).integers(0, 100, period) / div).astype(view.dtype)
    for o in range(0, view.size, period):
        n = min(period, view.size - o)
        view[o:o + n] = block[:n]


def time_device(ctx, call, steps, warmup, barrier):
    for _ in range(warmup):
        call()
    ctx.sync()
    barrier()
    l0 = ctx.launch_count()
    ctx.timer_start()
    for _ in range(steps):
        call()
    ms = ctx.timer_stop()
    ctx.sync()
    barrier()
    return ms, ctx.launch_count() - l0


def bench_kernel(blob, ctx, torch, kind, dtype, dim, steps, warmup, barrier=lambda: None, seed=1):
    """Device-resident timing of one workload on one GPU; returns dict."""
    a, b, c = device_inputs(torch, kind, dtype, dim, seed)
    torch.cuda.synchronize()
    if kind == "gemm":
        if dtype == "f32":
            ctx.sgemm_prepare(dim, dim, dim, dim, dim)
        call = lambda: ctx.gemm(dtype, dim, dim, dim, 1.0, a, dim, b, dim, 0.0, c, dim)
        flops, nbytes = gemm_flops(dim, dim, dim), gemm_bytes(dim, dim, dim, ELEM[dtype])
    else:
        call = lambda: ctx.gemv(dtype, dim, dim, 1.0, a, dim, b, 1, 0.0, c, 1)
        flops, nbytes = gemv_flops(dim, dim), gemv_bytes(dim, dim, ELEM[dtype])
    ms, launches = time_device(ctx, call, steps, warmup, barrier)
    kern = ctx.last_kernel()
    del a, b, c
    return {"ms_per_step": ms / steps, "gflops": flops * steps / (ms * 1e-3) * 1e-9,
            "gbs": nbytes * steps / (ms * 1e-3) * 1e-9, "launches": launches, "kernel": kern,
            "flops": flops, "bytes": nbytes}


def bench_e2e(blob, ctx, kind, dtype, dim, steps, warmup, barrier=lambda: None):
    """Host buffers -> C-ABI offload call -> host result, copies inside the timed region."""
    s = ELEM[dtype]
    md = blob.OFFLOAD_ALWAYS
    counts = (dim * dim, dim * dim, dim * dim) if kind == "gemm" else (dim * dim, dim, dim)
    ops = [ctx.alloc(md, cnt * s) for cnt in counts]
    views = [blob.host_view(h, cnt, dtype) for (h, _), cnt in zip(ops, counts)]
    host_fill(views[0], 7.0, 5)
    host_fill(views[1], 3.0, 6)
    views[2][:] = 0
    (hA, dA), (hB, dB), (hC, dC) = ops
    if kind == "gemm":
        if dtype == "f32":
            ctx.sgemm_prepare(dim, dim, dim, dim, dim)
        call = lambda: ctx.gemm_offload(dtype, dim, dim, dim, 1.0, hA, dA, dim, hB, dB, dim, 0.0, hC, dC, dim)
        flops = gemm_flops(dim, dim, dim)
        h2d, d2h = 2 * dim * dim * s, dim * dim * s
    else:
        call = lambda: ctx.gemv_offload(dtype, dim, dim, 1.0, hA, dA, dim, hB, dB, 0.0, hC, dC)
        flops = gemv_flops(dim, dim)
        h2d, d2h = (dim * dim + dim) * s, dim * s
    for _ in range(warmup):
        call()
    barrier()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()  # returns with the result valid on the host (b200_sync inside)
    dt = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    barrier()
    checksum = float(views[2][0]) + float(views[2][-1])
    for (h, d) in ops:
        ctx.free(md, h, d)
    return {"seconds": dt, "gflops": flops * steps / dt * 1e-9, "h2d": h2d, "d2h": d2h,
            "launches": launches, "checksum": checksum}


def kernel_roofline(ctx, peaks, peak_src, kind, dtype, gflops, gbs, kernel_name, flops, nbytes):
    """roofline object of one kernel: measured peak of the pipe that bounds it + the spec figure."""
    if kind == "gemm" and dtype == "f64":
        pk, pk2 = ctx.probe_peak("dmma"), ctx.probe_peak("dfma")  # GFLOP/s, measured live on this GPU
        peak = max(pk, pk2)
        r = {"bound": "tensor", "achieved": gflops / 1e3, "peak": peak / 1e3, "unit": "TFLOP/s",
             "frac": gflops / peak, "traffic": None, "kernel": kernel_name,
             "peak_source": "FP64 tensor (DMMA) pipe: b200_probe_peak(dmma) measured live on this GPU; "
                            "MEASURED_PEAKS.json holds no FP64 peak",
             "peak_spec": 37.2, "frac_of_spec": gflops / 37.2e3,
             "peak_spec_source": "148 SMs x 64 FP64 FMA/clk x 2 x 1.965 GHz (SURVEY.md section 6)",
             "probe_dfma_tflops": pk2 / 1e3, "probe_dmma_tflops": pk / 1e3}
    elif kind == "gemm":
        # 3xTF32: three tensor-core MMAs per FP32 product.  TF32 runs at half the bf16 rate, so the
        # measured denominator is MEASURED_PEAKS bf16_tflops / 2 / 3.
        ffma = ctx.probe_peak("ffma")
        tc_kernel = "tcgen05" in kernel_name
        peak = peaks["bf16_tflops"] * 1e3 / 2.0 / 3.0 if tc_kernel else ffma
        r = {"bound": "tensor" if tc_kernel else "fp32-pipe", "achieved": gflops / 1e3,
             "peak": peak / 1e3, "unit": "TFLOP/s", "frac": gflops / peak, "traffic": None, "kernel": kernel_name,
             "peak_source": (f"bf16_tflops of {peak_src} / 2 (TF32 = half the bf16 tensor rate) / 3 "
                             "(3xTF32: three MMAs per FP32 product)") if tc_kernel else
                            "b200_probe_peak(ffma) measured live on this GPU",
             "peak_spec": 375.0 if tc_kernel else 74.5, "frac_of_spec": gflops / (375e3 if tc_kernel else 74.5e3),
             "peak_spec_source": "2250 TFLOP/s dense bf16 / 2 / 3" if tc_kernel else "148 SMs x 128 FFMA/clk x 2 x 1.965 GHz",
             "probe_ffma_tflops": ffma / 1e3, "frac_of_ffma_peak": gflops / ffma}
    else:
        r = {"bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
             "frac": gbs / peaks["hbm_gbs"], "traffic": None, "kernel": kernel_name,
             "peak_source": f"hbm_gbs of {peak_src} (burst copy figure; a GEMV only reads)",
             "peak_spec": 7700.0, "frac_of_spec": gbs / 7700.0, "peak_spec_source": "HBM3e nominal 7.7 TB/s"}
    r["algorithmic_bytes"] = nbytes
    r["algorithmic_flops"] = flops
    return r


def e2e_path(kind, dtype, world):
    p, mv = prefix(dtype), ("m" if kind == "gemm" else "v")
    if world == 1:
        return f"b200_{p}gem{mv}_offload (Always mode of gpu::gem{mv}_gpu<T>), pinned host buffers"
    return (f"b200_mg_gem{mv} with B200_MG_UPLOAD|REPLICATE|COMPUTE|DOWNLOAD (the Always-mode step of gpu::gem{mv}_gpu<T> "
            "under B200_NGPUS, one rank per process): per rank and step, H2D of its share of "
            + ("A and its slab of B" if kind == "gemm" else "x and its row block of A") +
            " from pinned host buffers, shares pushed to every peer over NVLink while the kernels run, "
            "D2H of its slab of the result; b200_mg_sync")


def run_single(args, kind, dtype, dim, torch, blob, local):
    ctx = blob.Context(local)
    barrier = torch.cuda.synchronize
    a, b, c = device_inputs(torch, kind, dtype, dim, seed=1)
    torch.cuda.synchronize()
    if kind == "gemm":
        if dtype == "f32":
            ctx.sgemm_prepare(dim, dim, dim, dim, dim)
        call = lambda: ctx.gemm(dtype, dim, dim, dim, 1.0, a, dim, b, dim, 0.0, c, dim)
        flops, nbytes = gemm_flops(dim, dim, dim), gemm_bytes(dim, dim, dim, ELEM[dtype])
    else:
        call = lambda: ctx.gemv(dtype, dim, dim, 1.0, a, dim, b, 1, 0.0, c, 1)
        flops, nbytes = gemv_flops(dim, dim), gemv_bytes(dim, dim, ELEM[dtype])
    sampler = ClockSampler(local)
    sampler.start()
    ms, launches = time_device(ctx, call, args.steps, args.warmup, barrier)
    clocks = sampler.stop()
    kernel_name = ctx.last_kernel()
    value = flops * args.steps / (ms * 1e-3) * 1e-9
    gbs = nbytes * args.steps / (ms * 1e-3) * 1e-9
    del a, b, c
    torch.cuda.empty_cache()

    e2e_steps = max(1, min(args.steps, 5))
    e = bench_e2e(blob, ctx, kind, dtype, dim, e2e_steps, 1, barrier)

    peaks, peak_src = measured_peaks()
    roofline = kernel_roofline(ctx, peaks, peak_src, kind, dtype, value, gbs, kernel_name, flops, nbytes)
    traffic, traffic_src = ncu_traffic(args.workload, kernel_name) if dim == default_dim(kind) else (None, None)
    roofline["traffic"], roofline["traffic_source"] = traffic, traffic_src

    # --- the other three kernels at their headline sizes: value, roofline and e2e each --------
    extras = {}
    if not args.no_extras:
        for (k2, d2, n2) in (("gemm", "f64", 8192), ("gemm", "f32", 8192), ("gemv", "f64", 32768), ("gemv", "f32", 32768)):
            if (k2, d2, n2) == (kind, dtype, dim):
                continue
            key = f"{prefix(d2)}{k2}_{n2}"
            try:
                r = bench_kernel(blob, ctx, torch, k2, d2, n2, max(3, min(args.steps, 10)), 3)
                torch.cuda.empty_cache()
                e2 = bench_e2e(blob, ctx, k2, d2, n2, 3, 1)
                rf = kernel_roofline(ctx, peaks, peak_src, k2, d2, r["gflops"], r["gbs"], r["kernel"], r["flops"], r["bytes"])
                rf["traffic"], rf["traffic_source"] = ncu_traffic(prefix(d2) + k2, r["kernel"])
                extras[key] = {"gflops": round(r["gflops"], 1), "ms": round(r["ms_per_step"], 4), "kernel": r["kernel"],
                               "launches_per_step": r["launches"] / max(3, min(args.steps, 10)),
                               "roofline": rf,
                               "e2e": {"value": round(e2["gflops"], 1), "unit": "GFLOP/s", "h2d_bytes_per_step": e2["h2d"],
                                       "d2h_bytes_per_step": e2["d2h"], "path": e2e_path(k2, d2, 1)}}
            except Exception as ex:  # noqa: BLE001 -- explanatory only
                extras[key] = {"error": str(ex)[-300:]}
            torch.cuda.empty_cache()
        try:
            extras["probe_ffma_tflops"] = round(ctx.probe_peak("ffma") / 1e3, 2)
            extras["probe_copy_gbs"] = round(ctx.probe_peak("copy"), 1)
        except Exception as ex:  # noqa: BLE001
            extras["probe_error"] = str(ex)
    ctx.close()
    torch.cuda.empty_cache()

    # --- reported baselines, both outside every timed region ---------------------------------
    cpu = time_reference(kind, dtype, reference_dims(kind, dim), 5 if kind == "gemm" else 10, 0)
    gpu_base = cublas_baseline(kind, dtype, dim)

    line = {
        "metric": f"{prefix(dtype).upper()}{kind.upper()} GFLOP/s (GPU-BLOB formula)",
        "value": value, "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic", "config": workload_config(kind, dtype, dim, 1),
        "clocks": clocks,
        "e2e": {"value": e["gflops"], "unit": "GFLOP/s", "h2d_bytes_per_step": e["h2d"],
                "d2h_bytes_per_step": e["d2h"], "steps": e2e_steps, "path": e2e_path(kind, dtype, 1)},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu["value"], "unit": "GFLOP/s", "cores": cpu["cores"], "kind": cpu["kind"],
                         "sample": cpu["sample"]},
        "gpu_baseline": gpu_base,
        "kernel": kernel_name,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)


def run_team(args, kind, dtype, dim, torch, dist, blob, rank, world, local):
    """N > 1, strong scaling: one problem partitioned over the team (rank mode, CUDA IPC)."""
    import numpy as np
    s = ELEM[dtype]
    npdt = np.float64 if dtype == "f64" else np.float32

    def exchange(b):
        out = [None] * world
        dist.all_gather_object(out, b)
        return out

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    team = blob.Team(rank=rank, world=world, device=local)
    team.attach_flags(exchange)
    ctx = team.ctx()
    AL = blob.OFFLOAD_ALWAYS
    R, C_, G, U, D = blob.MG_REPLICATE, blob.MG_COMPUTE, blob.MG_GATHER, blob.MG_UPLOAD, blob.MG_DOWNLOAD

    if kind == "gemm":
        m = n = k = dim
        j0, nc = blob.mg_partition(n, world, rank, 128)
        k0, kk = blob.mg_partition(k, world, rank, 32)
        hA, dA_unused = ctx.alloc(AL, m * kk * s)        # host: my K-share of A
        hB, dB = ctx.alloc(AL, k * nc * s)               # my column slab of B
        hC, dC_slab = ctx.alloc(AL, m * nc * s)          # my column slab of C
        host_fill(blob.host_view(hA, m * kk, dtype), 7.0, 100 + rank)
        host_fill(blob.host_view(hB, k * nc, dtype), 3.0, 200 + rank)
        blob.host_view(hC, m * nc, dtype)[:] = 0
        dA = ctx.alloc_device(m * k * s)                 # my replica of all of A
        A_rep = team.map_peers(dA, exchange)
        C_full = ctx.alloc_device(m * n * s) if rank == 0 else None   # rank 0 gathers C
        C_root = team.map_root(C_full, exchange)                      # every rank maps rank 0's buffer
        dC = C_full + j0 * m * s if rank == 0 else dC_slab
        if dtype == "f32":
            ctx.sgemm_prepare(m, nc, k, m, k)
        # resident state for `value`: my share of A in my replica, my slab of B
        ctx.h2d(dA + k0 * m * s, hA, m * kk * s, 0)
        ctx.h2d(dB, hB, k * nc * s, 1)
        ctx.sync()
        flops, nbytes = gemm_flops(m, nc, k), gemm_bytes(m, nc, k, s)
        total_flops = gemm_flops(m, n, k)
        step_dev = lambda: team.gemm(rank, dtype, world, m, nc, k, 1.0, A_rep, m, dB, k, 0.0, dC, m, R | C_ | G,
                                     C_gather=C_root + j0 * m * s)
        step_e2e = lambda: (team.gemm(rank, dtype, world, m, nc, k, 1.0, A_rep, m, dB, k, 0.0, dC_slab, m, U | R | C_ | D,
                                      hA_share=hA, hB=hB, hC=hC), team.sync())
        h2d, d2h = (m * k + k * n) * s, m * n * s        # whole job per step: A once (1/N per rank), all of B; all of C
        exchange_bytes = (world - 1) * m * k * s + (m * n * s) * (world - 1) // world
    else:
        m = n = dim
        r0, rc = blob.mg_partition(m, world, rank, 32)
        x0, xn = blob.mg_partition(n, world, rank, 1)
        hA, dA = ctx.alloc(AL, rc * n * s)               # my compact row block of A (ld = rc)
        hx, dx_unused = ctx.alloc(AL, max(1, xn) * s)    # host: my share of x
        hy, dy_blk = ctx.alloc(AL, rc * s)
        host_fill(blob.host_view(hA, rc * n, dtype), 7.0, 100 + rank)
        host_fill(blob.host_view(hx, xn, dtype), 3.0, 200 + rank)
        dx = ctx.alloc_device(n * s)
        x_rep = team.map_peers(dx, exchange)
        y_full = ctx.alloc_device(m * s) if rank == 0 else None
        y_root = team.map_root(y_full, exchange)
        dy = y_full + r0 * s if rank == 0 else dy_blk
        ctx.h2d(dA, hA, rc * n * s, 0)
        ctx.h2d(dx + x0 * s, hx, xn * s, 1)
        ctx.sync()
        flops, nbytes = gemv_flops(rc, n), gemv_bytes(rc, n, s)
        total_flops = gemv_flops(m, n)
        step_dev = lambda: team.gemv(rank, dtype, world, rc, n, 1.0, dA, rc, x_rep, 0.0, dy, R | C_ | G,
                                     y_gather=y_root + r0 * s)
        step_e2e = lambda: (team.gemv(rank, dtype, world, rc, n, 1.0, dA, rc, x_rep, 0.0, dy_blk, U | R | C_ | D,
                                      hA=hA, h_lda=rc, hx_share=hx, hy=hy), team.sync())
        h2d, d2h = (m * n + n) * s, m * s
        exchange_bytes = (world - 1) * n * s + m * s * (world - 1) // world

    # ---- value: device resident, the exchange inside the step ---------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = time_device(ctx, step_dev, args.steps, args.warmup, barrier)
    team.sync()
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    kernel_name = ctx.last_kernel()
    value = total_flops * args.steps / (ms * 1e-3) * 1e-9
    per_gpu_gflops = flops * args.steps / (ms * 1e-3) * 1e-9
    per_gpu_gbs = nbytes * args.steps / (ms * 1e-3) * 1e-9

    # ---- e2e: host buffers, copies inside the timed region --------------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    step_e2e()
    barrier()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e_sec = max_over_ranks(time.perf_counter() - t0)
    e2e_launches = ctx.launch_count() - l0
    e2e_value = total_flops * e2e_steps / e_sec * 1e-9

    # ---- parity of the sharded result at FULL size: one sampled block of my slab, recomputed in
    # float64 from the host operands (rows of A are gathered from every rank's share)
    bs = 16
    if kind == "gemm":
        i0, c0 = m // 2 + 5, nc // 2
        mine = np.ascontiguousarray(blob.host_view(hA, m * kk, dtype).reshape(kk, m)[:, i0:i0 + bs]).astype(np.float64)
        a_rows = np.concatenate(exchange(mine), axis=0)   # (k, bs): rows i0.. of A, every rank's K-share in K order
        b_cols = blob.host_view(hB, k * nc, dtype).reshape(nc, k)[c0:c0 + bs].astype(np.float64)   # (bs, k)
        want = b_cols @ a_rows                            # [col][row] = the column-major block
        got = blob.host_view(hC, m * nc, dtype).reshape(nc, m)[c0:c0 + bs, i0:i0 + bs].astype(np.float64)
    else:
        xs = exchange(blob.host_view(hx, xn, dtype).astype(np.float64))
        x_all = np.concatenate(xs)
        i0 = rc // 2
        want = blob.host_view(hA, rc * n, dtype).reshape(n, rc)[:, i0:i0 + bs].astype(np.float64).T @ x_all
        got = blob.host_view(hy, rc, dtype)[i0:i0 + bs].astype(np.float64)
    err = float(np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)))
    errs = exchange(err)
    parity = {"max_rel_err_sampled_blocks": max(errs), "tolerance": TOL[dtype], "ok": bool(max(errs) <= TOL[dtype]),
              "what": f"one {bs}-wide block of every rank's slab of the e2e result vs a float64 recomputation from the host operands"}

    if rank == 0:
        peaks, peak_src = measured_peaks()
        roofline = kernel_roofline(ctx, peaks, peak_src, kind, dtype, per_gpu_gflops, per_gpu_gbs, kernel_name, flops, nbytes)
        roofline["note"] = "per GPU: this rank's slab of the problem over the step time (max over ranks), exchange included"
        line = {
            "metric": f"{prefix(dtype).upper()}{kind.upper()} GFLOP/s (GPU-BLOB formula)",
            "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": dtype, "data": "synthetic", "config": workload_config(kind, dtype, dim, world, True),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "ms_per_step": e_sec / e2e_steps * 1e3, "path": e2e_path(kind, dtype, world),
                    "gpu_launches": e2e_launches},
            "gpu_launches": launches,
            "exchange": {"nvlink_bytes_per_step": exchange_bytes, "inside_timed_step": True,
                         "how": "copy-engine pushes of each rank's share into every peer's replica + epoch flags "
                                "(csrc/mg.cu); the consumer's compute stream waits per chunk (cuStreamWaitValue32)"},
            "parity": parity,
            "roofline": roofline,
            "cpu_baseline": None,
            "kernel": kernel_name,
        }
        print(json.dumps(line), flush=True)
    barrier()
    team.close()


def run_weak(args, kind, dtype, dim, torch, dist, blob, rank, world, local):
    """N > 1, --weak: independent per-GPU problems (round 1's layout, no exchange in the step)."""
    ctx = blob.Context(local)

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    a, b, c = device_inputs(torch, kind, dtype, dim, seed=1 + rank)
    torch.cuda.synchronize()
    if kind == "gemm":
        call = lambda: ctx.gemm(dtype, dim, dim, dim, 1.0, a, dim, b, dim, 0.0, c, dim)
        flops = gemm_flops(dim, dim, dim)
    else:
        call = lambda: ctx.gemv(dtype, dim, dim, 1.0, a, dim, b, 1, 0.0, c, 1)
        flops = gemv_flops(dim, dim)
    ms, launches = time_device(ctx, call, args.steps, args.warmup, barrier)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    if rank == 0:
        print(json.dumps({
            "metric": f"{prefix(dtype).upper()}{kind.upper()} GFLOP/s (GPU-BLOB formula)",
            "value": flops * world * args.steps / (ms * 1e-3) * 1e-9, "unit": "GFLOP/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": dtype, "data": "synthetic",
            "config": workload_config(kind, dtype, dim, world, False), "gpu_launches": launches,
            "kernel": ctx.last_kernel()}), flush=True)
    ctx.close()


def run_own_arm(args, kind, dtype, dim, strong):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 backend has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
    if world == 1:
        run_single(args, kind, dtype, dim, torch, blob, local)
        return
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    try:
        if strong:
            run_team(args, kind, dtype, dim, torch, dist, blob, rank, world, local)
        else:
            run_weak(args, kind, dtype, dim, torch, dist, blob, rank, world, local)
        dist.barrier()
    finally:
        dist.destroy_process_group()


def default_dim(kind):
    return 8192 if kind == "gemm" else 32768


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dgemm", choices=["dgemm", "sgemm", "dgemv", "sgemv"])
    ap.add_argument("--dim", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--strong", action="store_true", help="(default at N > 1) partition ONE problem across the GPUs")
    ap.add_argument("--weak", action="store_true",
                    help="N > 1: independent per-GPU problems of the N = 1 size instead of BASELINE config 5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    kind = "gemm" if args.workload.endswith("gemm") else "gemv"
    dtype = "f64" if args.workload[0] == "d" else "f32"
    world = max(args.gpus, int(os.environ.get("WORLD_SIZE", "1")))
    strong = world > 1 and not args.weak
    dim = args.dim or (STRONG_DIM if strong else default_dim(kind))
    if args.impl == "reference":
        run_reference_arm(args, kind, dtype, dim, strong)
    else:
        run_own_arm(args, kind, dtype, dim, strong)


if __name__ == "__main__":
    main()
