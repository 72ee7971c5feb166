"""torchrun worker for the world_size-2 gloo test (CPU).  Each rank takes its
contiguous slice of the batch, shares and reconstructs it with the PRG counter
offset the sharding module prescribes, then the slices are all-gathered and
compared with the unsharded result.  The per-rank ENGINE here is the oracle (there
is no GPU in the CPU test run); on the GPU box the same driver code runs with the
libsclgpu Context (tests/test_gpu_parity.py::test_sharded_gpu_engine)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import __graft_entry__ as entry  # noqa: E402


def main():
    pkg = entry.load_package()
    o = entry.load_oracle()
    engine = o.PortOracle()
    sh = pkg.sharding
    rank, world, _ = sh.dist_env()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for field, t, n, N in [(61, 15, 32, 257), (127, 7, 16, 64)]:
        secrets = engine.vector_random(field, "secrets", 0, N)
        s = sh.shard_range(N, world, rank)
        first = sh.share_first_block(field, t, 3, s)
        shares = engine.shamir_share(field, secrets[s.lo:s.hi], t, n, "shamir bench", first)
        tam = shares.copy()
        if rank == 1 and t >= 1:
            tam.reshape(s.count, n, -1)[0, 0, 0] ^= np.uint64(1)
        rec = engine.recover_p(field, shares)
        _, err, nd = engine.recover_d(field, tam, t)
        g_sh = sh.gather_shards(torch.from_numpy(shares.view(np.int64)), s, N).numpy().view(np.uint64)
        g_rec = sh.gather_shards(torch.from_numpy(rec.view(np.int64)), s, N).numpy().view(np.uint64)
        total_bad = sh.sum_over_ranks(nd)
        full = engine.shamir_share(field, secrets, t, n, "shamir bench", 3)
        assert np.array_equal(g_sh, full), "gathered shares differ from the unsharded batch"
        assert np.array_equal(g_rec, secrets), "gathered secrets differ"
        assert total_bad == 1, total_bad
    # C5: rows of A sharded, x replicated, all-gather of the y slices (the path's one exchange step)
    rows, cols = 37, 20
    A = engine.vector_random(61, "mat A", 0, rows * cols).reshape(rows, cols)
    x = engine.vector_random(61, "vec x", 0, cols)
    s = sh.shard_range(rows, world, rank)
    y_local = engine.matvec(61, A[s.lo:s.hi], x) if s.count else np.zeros(0, dtype=np.uint64)
    y = sh.gather_shards(torch.from_numpy(y_local.view(np.int64)), s, rows).numpy().view(np.uint64)
    assert np.array_equal(y, engine.matvec(61, A, x)), "row-sharded mat-vec differs"
    dist.barrier()
    if rank == 0:
        print(f"DIST_OK world={world}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
