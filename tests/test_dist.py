"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: slab
arithmetic, A / x replication, ragged gathers.  The per-slab compute is injected
(numpy here; the B200 C ABI in bench.py), so this covers exactly the N > 1 path
that is not kernel code."""
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


def _partition():
    return importlib.import_module("gpu-blas-offload-benchmark_b200.partition")


def test_split_even_covers_everything():
    P = _partition()
    for total in (0, 1, 7, 8, 8192, 32769):
        for world in (1, 2, 3, 4, 8):
            slabs = P.all_slabs(total, world)
            assert slabs[0].start == 0
            assert sum(s.count for s in slabs) == total
            for a, b in zip(slabs, slabs[1:]):
                assert a.start + a.count == b.start
            assert max(s.count for s in slabs) - min(s.count for s in slabs) <= 1
    with pytest.raises(ValueError):
        P.split_even(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, m, n, k, out):
    sys.path.insert(0, ROOT)
    P = importlib.import_module("gpu-blas-offload-benchmark_b200.partition")
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same stream on every rank = the "host" inputs
        A = rng.integers(0, 100, (k, m)).astype(np.float64) / 7.0   # column-major A(m,k): [k][m]
        B = rng.integers(0, 100, (n, k)).astype(np.float64) / 3.0   # column-major B(k,n): [n][k]
        x = rng.integers(0, 100, n).astype(np.float64) / 3.0
        Ag = rng.integers(0, 100, (n, m)).astype(np.float64) / 7.0  # GEMV matrix, column-major [n][m]

        # ---- GEMM column slabs
        slab = P.column_slab(n, world, rank)
        A_t = torch.from_numpy(A.reshape(-1).copy()) if rank == 0 else torch.zeros(m * k, dtype=torch.float64)
        B_slab = torch.from_numpy(B[slab.start:slab.start + slab.count].reshape(-1).copy())

        def gemm(a, b, cols):
            a2 = a.numpy().reshape(k, m)           # [k][m]
            b2 = b.numpy().reshape(cols, k)        # [cols][k]
            return torch.from_numpy((b2 @ a2).reshape(-1).copy())  # [cols][m] = column-major slab

        _, C_full = P.gemm_column_slabs(dist, None, rank, world, m, n, k, A_t, B_slab, gemm)
        # ---- GEMV row blocks
        blk = P.row_block(m, world, rank)
        A_blk = P.compact_row_block(torch.from_numpy(Ag.reshape(-1).copy()), m, n, blk)
        x_t = torch.from_numpy(x.copy()) if rank == 0 else torch.zeros(n, dtype=torch.float64)

        def gemv(ab, xv, rows):
            return torch.from_numpy(ab.numpy().reshape(n, rows).T @ xv.numpy())

        _, y_full = P.gemv_row_blocks(dist, None, rank, world, m, n, A_blk, x_t, gemv)
        if rank == 0:
            want_C = (B @ A).reshape(-1)           # [n][m]
            want_y = Ag.T @ x
            ok = bool(np.allclose(C_full.numpy(), want_C, rtol=1e-13) and np.allclose(y_full.numpy(), want_y, rtol=1e-13))
            with open(out, "w") as f:
                f.write("ok" if ok else "mismatch")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape", [(2, (64, 48, 32)), (2, (33, 7, 19)), (3, (50, 10, 21))])
def test_slab_gemm_and_block_gemv_over_gloo(tmp_path, world, shape):
    m, n, k = shape
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, _free_port(), m, n, k, out), nprocs=world, join=True)
    assert open(out).read() == "ok"
