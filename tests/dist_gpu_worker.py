"""torchrun worker for the multi-GPU path on real GPUs (NCCL; one process per GPU).  Checks, against the
oracle on rank 0:
  * batch-sharded shamirSecretShare / shamirRecoverP with the PRG counter offset per rank, secrets gathered;
  * recoverD error counts summed over ranks;
  * batch-sharded sharing of math::Array<Fp61, 2> pairs (Pedersen's sharing step), shares gathered;
  * C5: row-sharded Fp61 mat-vec with the all-gather of the y slices (sharding.matvec_row_sharded);
  * shamirRecoverP with the all-gather fused into the kernel (stores to peer memory, sclgpu_ipc_*), every rank's copy;
  * the multi-device handle (sclgpu_mctx) driven by one process over all the GPUs.
Run by tests/test_gpu_parity.py::test_multi_gpu_nccl when at least two GPUs are visible."""
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
    port = o.PortOracle()
    sh = pkg.sharding
    B = pkg.binding
    rank, world, local_rank = sh.dist_env()
    # one GPU per rank under NCCL; with fewer GPUs than ranks (a one-GPU box) the ranks share devices and the
    # collectives go through gloo -- the per-rank PRG offsets, the peer-memory gather (cudaIpc between processes) and
    # the multi-device handle are exercised all the same
    ndev = torch.cuda.device_count()
    my_dev = local_rank % ndev
    torch.cuda.set_device(my_dev)
    if ndev >= world:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", my_dev))
    else:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = pkg.Context(my_dev)
    ctx.use_torch_stream()
    dev = torch.device("cuda", my_dev)

    # ---- batch-sharded share / reconstruct
    field, t, n, N = 61, 15, 32, 100003
    s = sh.shard_range(N, world, rank, align=2)
    d_sec = torch.empty(s.count, dtype=torch.int64, device=dev)
    ctx.random_dev(field, "secrets", sh.random_first_block(field, 0, s), s.count, d_sec)
    d_shr = torch.empty((n, s.count), dtype=torch.int64, device=dev)
    ctx.shamir_share_dev(field, d_sec, s.count, t, n, "shamir bench", sh.share_first_block(field, t, 3, s), d_shr, B.PARTY_MAJOR)
    if rank == world - 1:
        d_shr[20, 5] ^= 1                                        # one tampered (and checked) share on the last rank
    d_out = torch.empty(s.count, dtype=torch.int64, device=dev)
    d_err = torch.empty(s.count, dtype=torch.uint8, device=dev)
    nd = ctx.recover_d_dev(field, d_shr, s.count, n, t, d_out, d_err, B.PARTY_MAJOR)
    total_bad = sh.sum_over_ranks(nd, device=dev)
    ctx.recover_p_dev(field, d_shr[:, :], s.count, n, d_out, B.PARTY_MAJOR)
    torch.cuda.synchronize()
    g_sec = sh.gather_shards(d_sec.reshape(-1, 1), s, N, align=2).reshape(-1)
    g_shr = sh.gather_shards(d_shr.t().contiguous(), s, N, align=2)       # [N][n]
    if rank == 0:
        secrets = port.vector_random(field, "secrets", 0, N)
        assert np.array_equal(g_sec.cpu().numpy().view(np.uint64), secrets), "gathered secrets differ"
        want = port.shamir_share(field, secrets, t, n, "shamir bench", 3)
        got = g_shr.cpu().numpy().view(np.uint64)
        last = sh.shard_range(N, world, world - 1, align=2)
        want[last.lo + 5, 20] ^= np.uint64(1)
        assert np.array_equal(got, want), "gathered shares differ from the one-PRG batch"
        assert total_bad == 1, total_bad

    # ---- array-valued sharings (Pedersen's sharing step), batch-sharded the same way
    W, Na = 2, 30011
    sa = sh.shard_range(Na, world, rank)
    d_pairs = torch.empty((sa.count, W), dtype=torch.int64, device=dev)
    ctx.random_dev(61, "pairs", sa.lo * W // 2, sa.count * W, d_pairs)      # W even: every sharing starts on a block
    d_ash = torch.empty((sa.count, n, W), dtype=torch.int64, device=dev)
    ctx.shamir_share_array_dev(61, d_pairs, sa.count, W, t, n, "pedersen", sh.array_share_first_block(61, W, t, 5, sa),
                               d_ash, B.SECRET_MAJOR)
    d_arec = torch.empty((sa.count, W), dtype=torch.int64, device=dev)
    ctx.recover_p_array_dev(61, d_ash, sa.count, W, n, d_arec, B.SECRET_MAJOR)
    torch.cuda.synchronize()
    assert torch.equal(d_arec, d_pairs), "array sharing does not reconstruct on rank %d" % rank
    g_ash = sh.gather_shards(d_ash.reshape(sa.count, n * W), sa, Na)
    if rank == 0:
        pairs = port.vector_random(61, "pairs", 0, Na * W).reshape(Na, W)
        want = port.shamir_share_array(61, pairs, t, n, "pedersen", 5)
        assert np.array_equal(g_ash.cpu().numpy().view(np.uint64).reshape(want.shape), want), "gathered array shares differ"

    # ---- C5: row-sharded mat-vec + all-gather
    rows, cols = 2048, 1024
    rs = sh.shard_range(rows, world, rank)
    d_A = torch.empty((rs.count, cols), dtype=torch.int64, device=dev)
    ctx.random_dev(61, "mat A", (rs.lo * cols) // 2, rs.count * cols, d_A)   # rows [lo, hi) of Matrix::random(rows, cols)
    d_x = torch.empty(cols, dtype=torch.int64, device=dev)
    ctx.random_dev(61, "vec x", 0, cols, d_x)
    y = sh.matvec_row_sharded(ctx, 61, d_A, cols, d_x, rs, rows)
    if rank == 0:
        A = port.vector_random(61, "mat A", 0, rows * cols).reshape(rows, cols)
        x = port.vector_random(61, "vec x", 0, cols)
        assert np.array_equal(y.cpu().numpy().view(np.uint64), port.matvec(61, A, x)), "row-sharded mat-vec differs"
    # ---- shamirRecoverP fused with the all-gather of its result: stores to every rank's buffer over peer memory
    Ng = 40000
    sg = sh.shard_range(Ng, world, rank, align=2)
    d_gsec = torch.empty(sg.count, dtype=torch.int64, device=dev)
    ctx.random_dev(61, "secrets", sh.random_first_block(61, 0, sg), sg.count, d_gsec)
    d_gsh = torch.empty((n, sg.count), dtype=torch.int64, device=dev)
    ctx.shamir_share_dev(61, d_gsec, sg.count, t, n, "shamir bench", sh.share_first_block(61, t, 0, sg), d_gsh, B.PARTY_MAJOR)
    buf = ctx.malloc(8 * Ng)
    handles = [None] * world
    dist.all_gather_object(handles, ctx.ipc_export(buf))
    peers = [buf if r == rank else ctx.ipc_open(handles[r]) for r in range(world)]
    ctx.recover_p_gather_dev(d_gsh, sg.count, n, peers, sg.lo)
    torch.cuda.synchronize()
    dist.barrier()
    mine = torch.empty(Ng, dtype=torch.int64, device=dev)
    ctx.memcpy_d2d(mine.data_ptr(), buf, 8 * Ng)
    torch.cuda.synchronize()
    assert np.array_equal(mine.cpu().numpy().view(np.uint64), port.vector_random(61, "secrets", 0, Ng)), \
        "gathered reconstruction differs on rank %d" % rank
    dist.barrier()
    # the same gather from the single-launch step (k_share_recover61 with gather destinations)
    zero = torch.zeros(Ng, dtype=torch.int64, device=dev)
    ctx.memcpy_d2d(buf, zero.data_ptr(), 8 * Ng)
    torch.cuda.synchronize()
    dist.barrier()
    d_gsh2 = torch.empty((n, sg.count), dtype=torch.int64, device=dev)
    ctx.shamir_share_recover_gather_dev(d_gsec, sg.count, t, n, "shamir bench", sh.share_first_block(61, t, 77, sg), d_gsh2, peers,
                                        sg.lo, rec_shares=d_gsh)
    torch.cuda.synchronize()
    dist.barrier()
    ctx.memcpy_d2d(mine.data_ptr(), buf, 8 * Ng)
    torch.cuda.synchronize()
    assert np.array_equal(mine.cpu().numpy().view(np.uint64), port.vector_random(61, "secrets", 0, Ng)), \
        "gathered reconstruction (single-launch step) differs on rank %d" % rank
    dist.barrier()
    for r in range(world):
        if r != rank:
            ctx.ipc_close(peers[r])
    dist.barrier()
    ctx.free(buf)

    # ---- one process, one handle over all the GPUs (sclgpu_mctx): rank 0 drives, the others wait
    dist.barrier()
    if rank == 0:
        m = pkg.MultiContext([r % ndev for r in range(world)])
        Nm = 50001
        secrets = port.vector_random(61, "secrets", 0, Nm)
        assert np.array_equal(m.random("secrets", 0, Nm), secrets)
        shm = m.shamir_share(61, secrets, t, n, "shamir bench", 9)
        want = port.shamir_share(61, secrets, t, n, "shamir bench", 9)
        assert np.array_equal(shm, want), "multi-device share differs from the one-PRG batch"
        assert np.array_equal(m.recover_p(61, shm), secrets)
        s127 = port.vector_random(127, "secrets127", 0, 3001)
        sh127 = m.shamir_share(127, s127, 7, 16, "m127", 2)
        assert np.array_equal(sh127, port.shamir_share(127, s127, 7, 16, "m127", 2))
        sh127[17, 3, 0] ^= np.uint64(1)
        sh127[2900, 13, 0] ^= np.uint64(1)
        out, err, nd = m.recover_d(127, sh127, 7)
        w_out, w_err, w_nd = port.recover_d(127, sh127, 7)
        assert nd == w_nd == 2 and np.array_equal(err, w_err) and np.array_equal(out, w_out)
        m.close()
    dist.barrier()
    if rank == 0:
        print(f"DIST_GPU_OK world={world} gpus={min(ndev, world)} backend={dist.get_backend()}")
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
