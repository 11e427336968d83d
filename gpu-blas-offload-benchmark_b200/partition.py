"""Multi-GPU partitioning of the two kernels -- one process per GPU.

GEMM  C(m x n) = A(m x k) B(k x n): rank r owns a contiguous COLUMN SLAB of B and C
      (column-major => a pointer offset j0*ld, no repacking); A is replicated with a
      broadcast from rank 0 (NCCL over NVLink on GPUs).
GEMV  y(m) = A(m x n) x(n): rank r owns a ROW BLOCK of A and y; x is replicated.  In
      column-major a row block is strided (ld = m), so it is copied into a compact
      (ld = rows) slab when it is distributed.

Everything here is host logic over a torch.distributed process group plus an
injected `compute` callable, so the same code runs on NCCL with the B200 kernels
(bench.py) and on gloo/CPU in tests/test_dist.py.  No arithmetic lives here.
"""
from dataclasses import dataclass


@dataclass(frozen=True)
class Slab:
    start: int
    count: int


def split_even(total, world, rank):
    """Contiguous near-even split: the first (total % world) ranks get one extra."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return Slab(start, base + (1 if rank < extra else 0))


def column_slab(n, world, rank):
    return split_even(n, world, rank)


def row_block(m, world, rank):
    return split_even(m, world, rank)


def all_slabs(total, world):
    return [split_even(total, world, r) for r in range(world)]


def gemm_column_slabs(dist, group, rank, world, m, n, k, A, B_slab, compute, gather_to_root=True):
    """A: full (m*k,) tensor on every rank (valid on rank 0, overwritten elsewhere);
    B_slab: this rank's (k*cols,) column slab.  Returns (C_slab, C_full or None).
    `compute(A, B_slab, cols)` -> (m*cols,) tensor on the same device."""
    import torch
    slab = column_slab(n, world, rank)
    if world > 1:
        dist.broadcast(A, src=0, group=group)
    C_slab = compute(A, B_slab, slab.count)
    if not gather_to_root or world == 1:
        return C_slab, (C_slab if world == 1 else None)
    sizes = [s.count * m for s in all_slabs(n, world)]
    if rank == 0:
        parts = [torch.empty(sz, dtype=C_slab.dtype, device=C_slab.device) for sz in sizes]
    else:
        parts = None
    # column slabs are contiguous in column-major C: concatenation IS the gather.
    # Slab sizes may differ by one column, so this is send/recv, not dist.gather.
    if rank == 0:
        parts[0].copy_(C_slab)
        for r in range(1, world):
            dist.recv(parts[r], src=r, group=group)
    else:
        dist.send(C_slab, dst=0, group=group)
    return C_slab, (torch.cat(parts) if rank == 0 else None)


def gemv_row_blocks(dist, group, rank, world, m, n, A_block, x, compute, gather_to_root=True):
    """A_block: this rank's compact (rows*n,) row block (ld = rows); x: (n,) tensor
    (valid on rank 0).  Returns (y_block, y_full or None)."""
    import torch
    blk = row_block(m, world, rank)
    if world > 1:
        dist.broadcast(x, src=0, group=group)
    y_block = compute(A_block, x, blk.count)
    if not gather_to_root or world == 1:
        return y_block, (y_block if world == 1 else None)
    sizes = [b.count for b in all_slabs(m, world)]
    if rank == 0:
        parts = [torch.empty(sz, dtype=y_block.dtype, device=y_block.device) for sz in sizes]
        parts[0].copy_(y_block)
        for r in range(1, world):
            dist.recv(parts[r], src=r, group=group)
        return y_block, torch.cat(parts)
    dist.send(y_block, dst=0, group=group)
    return y_block, None


def compact_row_block(A_full, m, n, blk):
    """Column-major A (m x n, ld = m) -> compact row block (blk.count x n, ld = blk.count)."""
    return A_full.view(n, m)[:, blk.start:blk.start + blk.count].contiguous().view(-1)
