"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU.

The partitioning the GPU teams use is pure host arithmetic exported by the library
(b200_mg_partition, b200_mg_active_gemm / _gemv -- no CUDA call, so they run here).  Each
gloo rank plays one team member with exactly the library's split: it owns a K-share of A
(a block of A's columns) and a column slab of B and C, the shares are exchanged
(all_gather = what the NVLink pushes do on the GPUs) and the slab is accumulated chunk by
chunk in the arrival order csrc/mg.cu uses; GEMV does the same with row blocks of A and
shares of x.  The arithmetic is numpy; the assembled result is checked against the oracle."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _blob():
    return importlib.import_module("gpu-blas-offload-benchmark_b200")


def test_partition_covers_everything(blob):
    for total in (0, 1, 7, 8, 100, 8192, 32769):
        for parts in (1, 2, 3, 4, 8):
            for align in (1, 32, 128):
                pieces = [blob.mg_partition(total, parts, i, align) for i in range(parts)]
                assert pieces[0][0] == 0
                assert sum(c for _, c in pieces) == total
                for (s0, c0), (s1, _) in zip(pieces, pieces[1:]):
                    assert s0 + c0 == s1 and c0 >= 0
                if total >= 2 * align * parts:   # aligned boundaries, and nobody starves
                    assert all(s % align == 0 for s, _ in pieces)
                    assert min(c for _, c in pieces) >= align
    with pytest.raises(blob.B200Error):
        blob.mg_partition(10, 2, 2, 1)


def test_small_problems_stay_on_one_gpu(blob):
    """The spreading rule the harness sweep relies on (B200_NGPUS=8, sizes 1..8192)."""
    assert blob.mg_active_gemm(8, 8, 64, 64, 64) == 1
    assert blob.mg_active_gemm(8, 8, 1024, 1024, 1024) == 1
    assert blob.mg_active_gemm(8, 8, 4096, 4096, 4096) == 8
    assert blob.mg_active_gemm(8, 8, 32768, 32768, 32768) == 8
    assert blob.mg_active_gemm(8, 8, 8192, 600, 8192) == 2
    assert blob.mg_active_gemm(1, 8, 8192, 8192, 8192) == 1
    assert blob.mg_active_gemv(8, 8, 32768, 32768) == 8
    assert blob.mg_active_gemv(8, 8, 512, 512) == 1
    assert blob.mg_active_gemv(8, 4, 32768, 32) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, n, k, sub, out):
    sys.path.insert(0, ROOT)
    blob = _blob()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same stream on every rank = the "host" inputs
        A = rng.integers(0, 100, (k, m)).astype(np.float64) / 7.0   # column-major A(m,k): [k][m]
        B = rng.integers(0, 100, (n, k)).astype(np.float64) / 3.0   # column-major B(k,n): [n][k]
        x = rng.integers(0, 100, n).astype(np.float64) / 3.0
        Ag = rng.integers(0, 100, (n, m)).astype(np.float64) / 7.0  # GEMV matrix, column-major [n][m]

        # ---- GEMM: my K-share of A, my column slab of B; exchange shares; chunk-wise accumulation
        shares = [blob.mg_partition(k, world, p, 32) for p in range(world)]
        j0, nc = blob.mg_partition(n, world, rank, 128)
        k0, kk = shares[rank]
        kmax = max(c for _, c in shares)
        mine = np.zeros((kmax, m))
        mine[:kk] = A[k0:k0 + kk]                       # only my share ever leaves "my host"
        gathered = [torch.zeros(kmax, m, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(mine))
        replica = np.full((k, m), np.nan)
        C_slab = np.zeros((nc, m))
        for i in range(sub):                            # arrival order of csrc/mg.cu: own first, then the peer d places behind
            for d in range(world):
                p = (rank - d + world) % world
                s0, sk = shares[p]
                c0, ck = blob.mg_partition(sk, sub, i, 32)
                if ck == 0:
                    continue
                replica[s0 + c0:s0 + c0 + ck] = gathered[p].numpy()[c0:c0 + ck]
                C_slab += B[j0:j0 + nc, s0 + c0:s0 + c0 + ck] @ replica[s0 + c0:s0 + c0 + ck]
        assert not np.isnan(replica).any()              # the replica ends up complete
        sizes = [blob.mg_partition(n, world, p, 128)[1] for p in range(world)]
        pad = np.zeros((max(sizes), m))
        pad[:nc] = C_slab
        parts = [torch.zeros(max(sizes), m, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(parts, torch.from_numpy(pad))
        C_full = np.concatenate([parts[p].numpy()[:sizes[p]] for p in range(world)])   # column slabs are contiguous

        # ---- GEMV: my row block of A and y, my share of x
        r0, rc = blob.mg_partition(m, world, rank, 32)
        xs = [blob.mg_partition(n, world, p, 1) for p in range(world)]
        xmax = max(c for _, c in xs)
        xm = np.zeros(xmax)
        xm[:xs[rank][1]] = x[xs[rank][0]:xs[rank][0] + xs[rank][1]]
        xg = [torch.zeros(xmax, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(xg, torch.from_numpy(xm))
        x_rep = np.concatenate([xg[p].numpy()[:xs[p][1]] for p in range(world)])
        y_blk = Ag[:, r0:r0 + rc].T @ x_rep
        rows = [blob.mg_partition(m, world, p, 32)[1] for p in range(world)]
        ypad = np.zeros(max(rows))
        ypad[:rc] = y_blk
        yg = [torch.zeros(max(rows), dtype=torch.float64) for _ in range(world)]
        dist.all_gather(yg, torch.from_numpy(ypad))
        y_full = np.concatenate([yg[p].numpy()[:rows[p]] for p in range(world)])

        if rank == 0:
            import oracle_lib
            oracle = oracle_lib.load_oracle()
            want_C = np.zeros(m * n)
            oracle.gemm("f64", m, n, k, 1.0, A.reshape(-1), max(1, m), B.reshape(-1), max(1, k), 0.0, want_C, max(1, m))
            want_y = np.zeros(m)
            oracle.gemv("f64", m, n, 1.0, Ag.reshape(-1), max(1, m), x, 1, 0.0, want_y, 1)
            ok = bool(np.allclose(C_full.reshape(-1), want_C, rtol=1e-12) and np.allclose(x_rep, x) and
                      np.allclose(y_full, want_y, rtol=1e-12))
            with open(out, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,sub", [(2, (64, 600, 160), 2), (2, (33, 7, 19), 1), (3, (50, 1000, 333), 3)])
def test_team_schedule_over_gloo(tmp_path, blob, world, shape, sub):
    m, n, k = shape
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), m, n, k, sub, out), nprocs=world, join=True)
    assert open(out).read() == "ok"
