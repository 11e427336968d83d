#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 backend for GPU-BLOB.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
                    [--workload dgemm|sgemm|dgemv|sgemv] [--dim D] [--no-extras]

One "step" is one pass of the hot path over one batch of synthetic input: one
GEMM (or GEMV) call of the named workload.  Metric = GFLOP/s with the
reference's own FLOP formulas (include/doGemm.hh:493-505 -> 2MNK + MN;
include/doGemv.hh:377-388 -> 2MN + M), as BASELINE.json asks.

N = 1 default workload: square DGEMM M=N=K=8192, the top of BASELINE.json
configs[1] ("square DGEMM/SGEMM M=N=K 1->8192 on 1 B200").
N > 1 (torchrun, one rank per GPU): the same per-GPU problem as a column slab of
a wider one -- M=K=8192, N = 8192 x world_size; A replicated by an NCCL
broadcast over NVLink, B and C column slabs stay local (no data-path
collective inside a step), scaling = "weak".

value      device-resident: operands already in HBM, CUDA events on the
           library's compute stream around exactly K steps.
e2e        the same call through the reference-facing plugin path with HOST
           buffers: b200_?gemm_offload / b200_?gemv_offload (what
           gpu::gemm_gpu<T>::callGemm does in Always mode): pinned H2D of A and
           B, kernel, D2H of C every step.
roofline   achieved FLOP/s (or GB/s) of the dominant kernel vs the measured
           peak of the pipe that bounds it.
cpu_baseline / --impl reference
           the reference's own CPU path (cblas_dgemm from its OpenBLAS backend,
           oracle/_ref/libblobref.so) on this box's host cores.
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ELEM = {"f64": 8, "f32": 4}


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
    `ncu --set full` capture of this workload's bench command (profiles/r1_traffic.json,
    written by scripts/ncu_traffic.py from the .ncu-rep); None when the capture is of a
    different kernel than the one that just ran."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    if not os.path.exists(p):
        return None, None
    t = json.load(open(p)).get(workload)
    if not t or t.get("kernel_label") != kernel_name:
        return None, None
    return t["dram_read_bytes"] + t["dram_write_bytes"], t.get("source")


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


# ---------------------------------------------------------------- reference arm

def load_reference():
    so = os.path.join(ROOT, "oracle", "_ref", "libblobref.so")
    if not os.path.exists(so):
        subprocess.call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"],
                        stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    lib = ctypes.CDLL(so)
    lib.blobref_blas_config.restype = ctypes.c_char_p
    d, f, i, vp = ctypes.c_double, ctypes.c_float, ctypes.c_int, ctypes.c_void_p
    lib.blobref_cblas_dgemm.argtypes = [i, i, i, i, i, d, vp, i, vp, i, d, vp, i]
    lib.blobref_cblas_sgemm.argtypes = [i, i, i, i, i, f, vp, i, vp, i, f, vp, i]
    lib.blobref_cblas_dgemv.argtypes = [i, i, i, d, vp, i, vp, i, d, vp, i]
    lib.blobref_cblas_sgemv.argtypes = [i, i, i, f, vp, i, vp, i, f, vp, i]
    return lib


def reference_step_fn(lib, kind, dtype, dim):
    """One step of the reference CPU path: the cblas call its OpenBLAS backend
    issues (OpenBLAS/gemm.hh:27-33, OpenBLAS/gemv.hh:27-31) on resident inputs
    drawn from the reference's value distribution."""
    import numpy as np
    t = np.float64 if dtype == "f64" else np.float32
    rng = np.random.default_rng(19123005)
    if kind == "gemm":
        A = (rng.integers(0, 100, dim * dim) / 7.0).astype(t)
        B = (rng.integers(0, 100, dim * dim) / 3.0).astype(t)
        C = np.zeros(dim * dim, t)
        fn = lib.blobref_cblas_dgemm if dtype == "f64" else lib.blobref_cblas_sgemm
        return (lambda: fn(0, 0, dim, dim, dim, 1.0, A.ctypes.data, dim, B.ctypes.data, dim, 0.0,
                           C.ctypes.data, dim)), gemm_flops(dim, dim, dim), (A, B, C)
    A = (rng.integers(0, 100, dim * dim) / 7.0).astype(t)
    x = (rng.integers(0, 100, dim) / 3.0).astype(t)
    y = np.zeros(dim, t)
    fn = lib.blobref_cblas_dgemv if dtype == "f64" else lib.blobref_cblas_sgemv
    return (lambda: fn(0, dim, dim, 1.0, A.ctypes.data, dim, x.ctypes.data, 1, 0.0, y.ctypes.data, 1)), \
        gemv_flops(dim, dim), (A, x, y)


def cpu_sample_dim(kind, dim):
    # bounded sample: OpenBLAS GFLOP/s is flat above 4096
    return min(dim, 4096) if kind == "gemm" else min(dim, 16384)


def time_reference(kind, dtype, dim, steps, warmup):
    lib = load_reference()
    cores = os.cpu_count() or 1
    lib.blobref_blas_set_threads(min(cores, 64))  # OpenBLAS build: MAX_THREADS=64
    sdim = cpu_sample_dim(kind, dim)
    step, flops, keep = reference_step_fn(lib, kind, dtype, sdim)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": flops * steps / dt * 1e-9, "ms_per_step": dt / steps * 1e3,
            "cores": lib.blobref_blas_threads(), "kind": "reference",
            "sample": f"{'d' if dtype == 'f64' else 's'}{kind} M=N{'=K' if kind == 'gemm' else ''}={sdim}, "
                      f"{steps} calls of the reference's cblas_* (OpenBLAS) path; "
                      f"{lib.blobref_blas_config().decode().strip()}"}


def run_reference_arm(args, kind, dtype, dim):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = time_reference(kind, dtype, dim, args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": f"{'D' if dtype == 'f64' else 'S'}{kind.upper()} GFLOP/s (GPU-BLOB formula)",
        "value": r["value"], "unit": "GFLOP/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if (args.strong and args.gpus > 1) else "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic",
        "config": workload_config(kind, dtype, dim, args.gpus, args.strong),
        "cpu_baseline": {"value": r["value"], "unit": "GFLOP/s", "cores": r["cores"],
                         "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------- own arm

def workload_config(kind, dtype, dim, gpus, strong=False):
    p = "d" if dtype == "f64" else "s"
    if gpus > 1 and strong:
        wl = (f"{p}gemm M=N=K={dim} in {gpus} column slabs ({dim // gpus} columns of B and C per GPU, A replicated)"
              if kind == "gemm" else
              f"{p}gemv M=N={dim} in {gpus} row blocks ({dim // gpus} rows of A and y per GPU, x replicated)")
    elif kind == "gemm":
        wl = f"{p}gemm_square M=N=K={dim}" if gpus == 1 else \
            f"{p}gemm column slabs: M=K={dim}, N={dim}x{gpus} ({dim} columns of B and C per GPU, A replicated)"
    else:
        wl = f"{p}gemv_square M=N={dim}" if gpus == 1 else \
            f"{p}gemv row blocks: N={dim}, M={dim}x{gpus} ({dim} rows of A and y per GPU, x replicated)"
    return {"workload": wl, "baseline_config": "configs[1]" if kind == "gemm" else "configs[3]",
            "alpha": 1, "beta": 0, "layout": "column-major, NoTrans",
            "l2": "operands larger than the 126 MB L2 (no flush needed)",
            "inputs": "rand()%100 / 7.0 and / 3.0 value distribution of the reference generator"}


def device_inputs(torch, kind, dtype, dim, seed, loc=None):
    """loc = this rank's extent of the partitioned dimension (columns of B/C for GEMM,
    rows of A/y for GEMV); dim when there is one rank or the scaling is weak."""
    loc = loc or dim
    tdt = torch.float64 if dtype == "f64" else torch.float32
    g = torch.Generator(device="cuda").manual_seed(seed)

    def fill(count, div):
        return torch.randint(0, 100, (count,), generator=g, device="cuda").to(tdt) / div
    if kind == "gemm":
        return fill(dim * dim, 7.0), fill(dim * loc, 3.0), torch.zeros(dim * loc, dtype=tdt, device="cuda")
    return fill(loc * dim, 7.0), fill(dim, 3.0), torch.zeros(loc, dtype=tdt, device="cuda")


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
    """Device-resident timing of one workload; returns dict."""
    a, b, c = device_inputs(torch, kind, dtype, dim, seed)
    torch.cuda.synchronize()
    if kind == "gemm":
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
    import numpy as np
    s = ELEM[dtype]
    md = blob.OFFLOAD_ALWAYS
    if kind == "gemm":
        counts = (dim * dim, dim * dim, dim * dim)
    else:
        counts = (dim * dim, dim, dim)
    ops = [ctx.alloc(md, cnt * s) for cnt in counts]
    views = [blob.host_view(h, cnt, dtype) for (h, _), cnt in zip(ops, counts)]
    rng = np.random.default_rng(5)
    chunk = 1 << 24
    for v, div in zip(views[:2], (7.0, 3.0)):
        for o in range(0, v.size, chunk):
            v[o:o + chunk] = rng.integers(0, 100, min(chunk, v.size - o)) / div
    views[2][:] = 0
    (hA, dA), (hB, dB), (hC, dC) = ops
    if kind == "gemm":
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


def bench_e2e_multi(blob, ctx, torch, dist, kind, dtype, dim, steps, warmup, barrier, rank, world, loc):
    """N > 1 end to end.  Every step: each rank uploads its 1/N share of the replicated
    operand (A for GEMM, x for GEMV) over its OWN PCIe link, an NCCL all-gather over
    NVLink replicates it on every GPU, and each rank uploads its slab (B columns / A
    rows), runs the kernel and downloads its slab of the result (C columns / y rows) to
    its own pinned host buffer."""
    import numpy as np
    s = ELEM[dtype]
    tdt = torch.float64 if dtype == "f64" else torch.float32
    md = blob.OFFLOAD_ONCE
    rep_count = dim * dim if kind == "gemm" else dim          # A / x
    share = (rep_count + world - 1) // world                  # this rank's piece of it
    slab_count = dim * loc                                    # B slab / A row block
    out_count = dim * loc if kind == "gemm" else loc          # C slab / y block
    h_share, _d = ctx.alloc(md, share * s)
    h_slab, d_slab = ctx.alloc(md, slab_count * s)
    h_out, d_out = ctx.alloc(md, out_count * s)
    rep_t = torch.zeros(share * world, dtype=tdt, device="cuda")   # NCCL needs torch tensors
    share_t = torch.zeros(share, dtype=tdt, device="cuda")
    rng = np.random.default_rng(7 + rank)
    for h, cnt, div in ((h_slab, slab_count, 3.0 if kind == "gemm" else 7.0),
                        (h_share, share, 7.0 if kind == "gemm" else 3.0)):
        v = blob.host_view(h, cnt, dtype)
        for o in range(0, v.size, 1 << 24):
            v[o:o + (1 << 24)] = rng.integers(0, 100, min(1 << 24, v.size - o)) / div

    def step():
        ctx.h2d(share_t, h_share, share * s, 0)
        ctx.h2d(d_slab, h_slab, slab_count * s, 1)
        ctx.sync()
        dist.all_gather_into_tensor(rep_t, share_t)
        torch.cuda.synchronize()
        if kind == "gemm":
            ctx.gemm(dtype, dim, loc, dim, 1.0, rep_t, dim, d_slab, dim, 0.0, d_out, dim)
        else:
            ctx.gemv(dtype, loc, dim, 1.0, d_slab, loc, rep_t, 1, 0.0, d_out, 1)
        ctx.d2h(h_out, d_out, out_count * s, 2)
        ctx.sync()

    for _ in range(warmup):
        step()
    barrier()
    l0 = ctx.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    barrier()
    dt = time.perf_counter() - t0
    launches = ctx.launch_count() - l0
    ctx.free(md, h_share, _d)
    ctx.free(md, h_slab, d_slab)
    ctx.free(md, h_out, d_out)
    h2d = (share + slab_count) * world * s       # whole job: the replicated operand once + every slab
    return {"seconds": dt, "h2d": h2d, "d2h": out_count * s * world, "launches": launches}


def run_own_arm(args, kind, dtype, dim):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 backend has no CPU path "
                         "(use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    blob = importlib.import_module("gpu-blas-offload-benchmark_b200")
    ctx = blob.Context(local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # --- headline: device resident ---------------------------------------
    loc = dim // world if (args.strong and world > 1) else dim
    a, b, c = device_inputs(torch, kind, dtype, dim, seed=1 + rank if kind == "gemv" else 1, loc=loc)
    if world > 1:
        # the replicated operand travels over NVLink once (A for GEMM, x for GEMV);
        # the slab operands (B, C columns / A, y rows) are generated per rank
        rep = a if kind == "gemm" else b
        dist.broadcast(rep, src=0)
        if kind == "gemm":
            g = torch.Generator(device="cuda").manual_seed(100 + rank)
            b = torch.randint(0, 100, (dim * loc,), generator=g, device="cuda").to(b.dtype) / 3.0
    torch.cuda.synchronize()
    if kind == "gemm":
        call = lambda: ctx.gemm(dtype, dim, loc, dim, 1.0, a, dim, b, dim, 0.0, c, dim)
        flops, nbytes = gemm_flops(dim, loc, dim), gemm_bytes(dim, loc, dim, ELEM[dtype])
    else:
        call = lambda: ctx.gemv(dtype, loc, dim, 1.0, a, loc, b, 1, 0.0, c, 1)
        flops, nbytes = gemv_flops(loc, dim), gemv_bytes(loc, dim, ELEM[dtype])

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = time_device(ctx, call, args.steps, args.warmup, barrier)
    clocks = sampler.stop() if rank == 0 else None
    ms = max_over_ranks(ms)
    kernel_name = ctx.last_kernel()
    value = flops * world * args.steps / (ms * 1e-3) * 1e-9
    per_gpu_gflops = flops * args.steps / (ms * 1e-3) * 1e-9
    per_gpu_gbs = nbytes * args.steps / (ms * 1e-3) * 1e-9
    del a, b, c
    torch.cuda.empty_cache()

    # --- e2e through the plugin path with host buffers ---------------------
    e2e_steps = max(1, min(args.steps, 5))
    if world == 1:
        e = bench_e2e(blob, ctx, kind, dtype, dim, e2e_steps, 1, barrier)
    else:
        e = bench_e2e_multi(blob, ctx, torch, dist, kind, dtype, dim, e2e_steps, 1, barrier, rank, world, loc)
    e_sec = max_over_ranks(e["seconds"])
    e2e_value = flops * world * e2e_steps / e_sec * 1e-9

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # --- roofline of the dominant kernel -----------------------------------
    peaks, peak_src = measured_peaks()
    if kind == "gemm" and dtype == "f64":
        pk = ctx.probe_peak("dmma")  # GFLOP/s, measured live on this GPU
        pk2 = ctx.probe_peak("dfma")
        peak = max(pk, pk2)
        roofline = {"bound": "tensor", "achieved": per_gpu_gflops / 1e3, "peak": peak / 1e3, "unit": "TFLOP/s",
                    "frac": per_gpu_gflops / peak, "traffic": None, "kernel": kernel_name,
                    "peak_source": "FP64 tensor (DMMA) pipe: b200_probe_peak(dmma) measured live on this GPU; "
                                   "MEASURED_PEAKS.json holds no FP64 peak",
                    "probe_dfma_tflops": pk2 / 1e3, "probe_dmma_tflops": pk / 1e3}
    elif kind == "gemm":
        # 3xTF32: three tensor-core MMAs per FP32 product.  TF32 runs at half the bf16 rate, so the
        # measured denominator is MEASURED_PEAKS bf16_tflops / 2 / 3.
        ffma = ctx.probe_peak("ffma")
        tc_kernel = "tcgen05" in kernel_name
        peak = peaks["bf16_tflops"] * 1e3 / 2.0 / 3.0 if tc_kernel else ffma
        roofline = {"bound": "tensor" if tc_kernel else "fp32-pipe", "achieved": per_gpu_gflops / 1e3,
                    "peak": peak / 1e3, "unit": "TFLOP/s", "frac": per_gpu_gflops / peak, "traffic": None,
                    "kernel": kernel_name,
                    "peak_source": (f"bf16_tflops of {peak_src} / 2 (TF32 = half the bf16 tensor rate) / 3 "
                                    "(3xTF32: three MMAs per FP32 product)") if tc_kernel else
                                   "b200_probe_peak(ffma) measured live on this GPU",
                    "probe_ffma_tflops": ffma / 1e3, "frac_of_ffma_peak": per_gpu_gflops / ffma}
    else:
        roofline = {"bound": "hbm", "achieved": per_gpu_gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": per_gpu_gbs / peaks["hbm_gbs"], "traffic": None, "kernel": kernel_name,
                    "peak_source": f"hbm_gbs of {peak_src} (burst copy figure)"}

    traffic, traffic_src = ncu_traffic(args.workload, kernel_name) if (world == 1 and dim == default_dim(kind)) else (None, None)
    roofline["traffic"] = traffic
    roofline["traffic_source"] = traffic_src
    roofline["algorithmic_bytes"] = nbytes
    roofline["algorithmic_flops"] = flops

    # --- the other three kernels at their headline sizes (explanatory) ------
    extras = {}
    if not args.no_extras and world == 1:
        for (k2, d2, n2) in (("gemm", "f64", 8192), ("gemm", "f32", 8192), ("gemv", "f64", 32768), ("gemv", "f32", 32768)):
            if (k2, d2, n2) == (kind, dtype, dim):
                continue
            r = bench_kernel(blob, ctx, torch, k2, d2, n2, max(3, min(args.steps, 10)), 3)
            key = f"{'d' if d2 == 'f64' else 's'}{k2}_{n2}"
            extras[key] = {"gflops": round(r["gflops"], 1), "ms": round(r["ms_per_step"], 4), "kernel": r["kernel"]}
            if k2 == "gemv":
                extras[key]["gbs"] = round(r["gbs"], 1)
                extras[key]["hbm_frac"] = round(r["gbs"] / peaks["hbm_gbs"], 3)
            torch.cuda.empty_cache()
        try:
            extras["probe_ffma_tflops"] = round(ctx.probe_peak("ffma") / 1e3, 2)
            extras["probe_copy_gbs"] = round(ctx.probe_peak("copy"), 1)
        except Exception as ex:  # noqa: BLE001 -- explanatory only
            extras["probe_error"] = str(ex)

    # --- CPU baseline: the reference's OpenBLAS path on this box's cores -----
    cpu = time_reference(kind, dtype, dim, 3, 1)

    line = {
        "metric": f"{'D' if dtype == 'f64' else 'S'}{kind.upper()} GFLOP/s (GPU-BLOB formula)",
        "value": value, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if (args.strong and world > 1) else "weak", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic", "config": workload_config(kind, dtype, dim, world, args.strong),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "GFLOP/s", "h2d_bytes_per_step": e["h2d"],
                "d2h_bytes_per_step": e["d2h"], "steps": e2e_steps,
                "path": ("b200_%sgem%s_offload (Always mode of gpu::gem%s_gpu<T>), pinned host buffers"
                         % ("d" if dtype == "f64" else "s", "m" if kind == "gemm" else "v", "m" if kind == "gemm" else "v"))
                if world == 1 else
                ("per rank and step: b200_h2d_async of its 1/N share of the replicated operand + its slab from "
                 "pinned host buffers, NCCL all-gather over NVLink, b200_%sgem%s, b200_d2h_async of its result slab"
                 % ("d" if dtype == "f64" else "s", "m" if kind == "gemm" else "v"))},
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": {"value": cpu["value"], "unit": "GFLOP/s", "cores": cpu["cores"], "kind": cpu["kind"],
                         "sample": cpu["sample"]},
        "kernel": kernel_name,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.barrier()
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
    ap.add_argument("--strong", action="store_true",
                    help="N > 1: partition ONE dim^3 (dim^2) problem across the GPUs (BASELINE config 5) instead of weak scaling")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    kind = "gemm" if args.workload.endswith("gemm") else "gemv"
    dtype = "f64" if args.workload[0] == "d" else "f32"
    dim = args.dim or default_dim(kind)
    if args.impl == "reference":
        run_reference_arm(args, kind, dtype, dim)
    else:
        run_own_arm(args, kind, dtype, dim)


if __name__ == "__main__":
    main()
