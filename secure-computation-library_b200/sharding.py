"""Batch-of-secrets sharding over the GPUs of one box.

Units (secrets, vector elements, matrix rows) are independent, so the index
range [0, N) is cut into `world` contiguous slices and rank r works on slice r
with no data-path collective.  The only coupling is the PRG: SCL would draw
secret j's coefficients when its PRG counter is first_block + j*B
(B = blocks per shamirSecretShare call, shamir.h:56 + prg.cc:129-133), and
AES-CTR is seekable, so rank r simply starts at first_block + lo*B with the
same seed.  Collectives (torch.distributed: NCCL on GPUs, gloo in the CPU
tests) are used only to gather results / OR the recoverD error count.
"""
from __future__ import annotations

from dataclasses import dataclass


def blocks_per_share_call(field: int, t: int) -> int:
    """ceil((t+1)*byteSize/16) keystream blocks per shamirSecretShare call."""
    bs = 8 if field == 61 else 16
    return ((t + 1) * bs + 15) // 16


def blocks_for_random(field: int, n: int, one_per_block: bool = False) -> int:
    """Blocks Vector::random(n) (vector.h:508-519) or n x FF::random (ff.h:72-76) consume."""
    if one_per_block:
        return n
    bs = 8 if field == 61 else 16
    return (n * bs + 15) // 16


@dataclass(frozen=True)
class Shard:
    rank: int
    world: int
    lo: int  # first unit of this rank
    hi: int  # one past the last unit

    @property
    def count(self) -> int:
        return self.hi - self.lo


def shard_range(n_units: int, world: int, rank: int, align: int = 1) -> Shard:
    """Contiguous slice of [0, n_units) for `rank`; boundaries are multiples of
    `align` (e.g. 2 for Fp61 Vector::random so every rank starts on a block)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad world/rank")
    if n_units < 0 or align < 1:
        raise ValueError("bad n_units/align")
    groups = (n_units + align - 1) // align
    lo = min((groups * rank // world) * align, n_units)
    hi = min((groups * (rank + 1) // world) * align, n_units)
    return Shard(rank, world, lo, hi)


def share_first_block(field: int, t: int, first_block: int, shard: Shard) -> int:
    """PRG counter at which rank `shard.rank` starts its slice of a batch share."""
    return first_block + shard.lo * blocks_per_share_call(field, t)


def array_share_first_block(field: int, width: int, t: int, first_block: int, shard: Shard) -> int:
    """PRG counter at which a rank starts its slice of a batch of array-valued sharings
    (shamirSecretShare on math::Array<FF, width>: ceil((t+1)*width*byteSize/16) blocks per sharing)."""
    bs = 8 if field == 61 else 16
    return first_block + shard.lo * (((t + 1) * width * bs + 15) // 16)


def additive_share_first_block(n: int, first_block: int, shard: Shard) -> int:
    """PRG counter for a rank's slice of a batch additiveShare: n-1 FF::random draws (one block each) per secret."""
    return first_block + shard.lo * (n - 1)


def random_first_block(field: int, first_block: int, shard: Shard, one_per_block: bool = False) -> int:
    """PRG counter for a rank's slice of Vector::random / FF::random x n.
    For Fp61 Vector::random the slice must start on an even element (align=2)."""
    if one_per_block or field == 127:
        return first_block + shard.lo
    if shard.lo % 2:
        raise ValueError("Fp61 Vector::random shards must start on an even element")
    return first_block + shard.lo // 2


def dist_env():
    """(rank, world, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    import os

    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def gather_shards(local, shard: Shard, n_units: int, group=None, align: int = 1):
    """all_gather the per-rank result slices (torch tensors, dim 0 = units) into the
    full [n_units, ...] tensor on every rank.  Uneven slices are padded.  `align` must be the
    value the slices were cut with (shard_range)."""
    import torch
    import torch.distributed as dist

    if shard.world == 1:
        return local
    counts = [shard_range(n_units, shard.world, r, align).count for r in range(shard.world)]
    pad = max(counts)
    # gloo (several ranks sharing one GPU, or the CPU tests) moves host tensors; NCCL device tensors
    via_host = local.is_cuda and dist.get_backend(group) == "gloo"
    src = local.cpu() if via_host else local
    buf = torch.zeros((pad,) + tuple(src.shape[1:]), dtype=src.dtype, device=src.device)
    buf[: src.shape[0]] = src
    out = [torch.empty_like(buf) for _ in range(shard.world)]
    dist.all_gather(out, buf, group=group)
    res = torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)
    return res.to(local.device) if via_host else res


def sum_over_ranks(value: int, device=None, group=None) -> int:
    """SUM all-reduce of one integer (recoverD: number of flagged secrets)."""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return int(value)
    if dist.get_backend(group) == "gloo":
        device = None
    t = torch.tensor([int(value)], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def matvec_row_sharded(ctx, field: int, d_A_local, cols: int, d_x, shard: Shard, n_rows: int, group=None):
    """y = A x (Matrix::multiply(Vector), matrix.h:498-513) with the ROWS of A sharded over the ranks
    (BASELINE config C5): every rank multiplies its contiguous row slice on its own GPU
    (`d_A_local` = rows [shard.lo, shard.hi) of A, row-major, `d_x` replicated), then the y slices are
    all-gathered -- the one real exchange step of the path (8 KiB per GPU at 8192 rows over 8 GPUs).
    Returns the full y (n_rows elements) as a device tensor on every rank."""
    import torch

    w = 1 if field == 61 else 2
    y_local = torch.empty((shard.count, w), dtype=torch.int64, device=d_x.device)
    if shard.count:
        ctx.matvec_dev(field, d_A_local, shard.count, cols, d_x, y_local)
    torch.cuda.current_stream().synchronize() if d_x.is_cuda else None
    y = gather_shards(y_local, shard, n_rows, group=group)
    return y.reshape(n_rows) if w == 1 else y
