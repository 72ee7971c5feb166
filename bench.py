#!/usr/bin/env python
"""bench.py -- Fp61 Shamir share + reconstruct throughput (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic secrets:
    shares  = shamirSecretShare(secret_j, t, n, prg)   for all j   (PRG fused in)
    secret' = shamirRecoverP(shares_j)                 for all j
Workload at N=1: BASELINE configs[1] -- Mersenne-61, n=32, t=15, 2^26 secrets.

  value   : secrets/s, secrets resident in HBM, shares written to and read back from HBM (party-major planes),
            CUDA events, max over ranks.  Schedule: ONE persistent launch per step (k_share_recover61): four share
            groups produce batch k while a reconstruction group of the same CTAs reconstructs batch k-1 from the other
            plane buffer, both on the tensor cores (the share side is bound by the SM's shared-memory/ALU pipes, the
            reconstruction by HBM).
            K shares and K reconstructions run inside the timed region.  `schedules` holds the same step as two
            kernels back to back on one stream, as two kernels on two streams, and as one launch that reconstructs
            the very tiles it stores; each verified.
  scaling : "weak" (2^26 secrets per GPU, the default `value`) and, in `strong`, BASELINE configs[1] as worded:
            2^26 secrets IN TOTAL, batch-sharded over the ranks.  `gathered` adds the all-gather of the
            reconstructed secrets inside the timing: fused into the reconstruction kernel (stores to peer memory
            over NVLink) and, beside it, NCCL's all_gather after the kernel.
  verified_vs_oracle_all_ranks : every rank compares a prefix and two far slices of its share planes (both plane
            buffers) with the seekable checker at ITS OWN PRG offset, and the reconstructed secrets with its slice of
            the secret stream; the AND over the ranks is reported.
  e2e     : the same metric through the reference-facing C-ABI host entry points with HOST (pinned) buffers in SCL's
            [N][n] layout, H2D/D2H copies inside the timing; full duplex: share_async(batch k) while
            recover_p(batch k-1).
  configs : BASELINE's other configs (C1, C3, C4, C5), each timed on the device with its own roofline figure and
            checked against the oracle (bounded to a few seconds each).
  --impl reference : SCL's own CPU code (oracle/_ref, built from the unmodified reference sources) on all host
            cores, a bounded sample per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
import __graft_entry__ as entry  # noqa: E402

METRIC = "fp61_shamir_share_reconstruct_secrets_per_s"
UNIT = "secrets/s"
FIELD, T, NPARTIES = 61, 15, 32
ALGO_IMADS_PER_SECRET = 2048       # 512 field muls x 4 32-bit IMADs
AES_LDS_PER_SECRET = 1083          # table lookups per secret of the fused AES-128-CTR (ncu op mix, profiles/)
ALGO_BYTES_SHARE = 8 + 8 * NPARTIES    # secret in, n shares out
ALGO_BYTES_RECOVER = 8 * NPARTIES + 8  # n shares in, secret out
SEED_SHARE, SEED_SECRETS = "shamir bench", "secrets"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-secrets", type=int, default=26, help="2^k secrets per GPU (weak) / in total (strong); 26 = BASELINE configs[1]")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"], help="which curve `value` reports (the other one is in the line too)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-staged", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--no-gathered", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_config(args, world):
    per_gpu = (1 << args.log2_secrets) if args.scaling == "weak" else (1 << args.log2_secrets) // world
    return {
        "workload": f"Mersenne-61 Shamir share+recoverP n={NPARTIES} t={T}, 2^{args.log2_secrets} secrets "
                    + ("per GPU" if args.scaling == "weak" else "in total") + " (BASELINE configs[1])",
        "field": "Fp61", "n": NPARTIES, "t": T, "secrets_per_gpu": per_gpu, "secrets_total": per_gpu * world,
        "sharding": f"batch x{world}, contiguous slices, PRG counter offset per rank, no data-path collective",
        "prg": "AES-128-CTR fused into the share kernel (seed 'shamir bench'): coefficients are drawn inside the timed region",
        "schedule": "one persistent launch per step (k_share_recover61): share(batch k) by the share groups, recoverP(batch k-1) by "
                    "reconstruction warps of the same CTAs, share planes double-buffered; K shares + K reconstructions inside the "
                    "timed region",
        "l2": "inputs larger than L2 (share planes 8*n*N bytes >> 126 MB); no explicit flush",
    }


# ----------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------- CPU arms
def cpu_reference_run(seconds: float, threads: int | None = None):
    """SCL's own share+recoverP call sequence on the host cores, bounded sample.
    Returns (secrets_per_s, kind, cores, n_sample, build_flags)."""
    o = entry.load_oracle()
    orc = o.best_oracle()
    cores = threads or (os.cpu_count() or 1)
    probe = 64 * cores
    dt = orc.bench_share_recover(FIELD, probe, T, NPARTIES, 0, cores)
    if dt <= 0:
        raise RuntimeError("CPU baseline produced wrong secrets")
    rate = probe / dt
    n_sample = max(probe, int(rate * seconds))
    dt = orc.bench_share_recover(FIELD, n_sample, T, NPARTIES, 0, cores)
    if dt <= 0:
        raise RuntimeError("CPU baseline produced wrong secrets")
    return n_sample / dt, orc.kind, cores, n_sample, getattr(orc, "build_flags", "plain C port, gcc -O3")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(1.0, min(10.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    kind, cores, n_sample, flags = "port", 1, 0, ""
    for i in range(args.warmup + args.steps):
        v, kind, cores, n_sample, flags = cpu_reference_run(per_step)
        if i >= args.warmup:
            vals.append((v, n_sample))
    tot_secrets = sum(n for _, n in vals)
    tot_time = sum(n / v for v, n in vals)
    value = tot_secrets / tot_time
    one, _, _, n_one, _ = cpu_reference_run(min(per_step, 4.0), threads=1)
    sample = (f"{n_sample} secrets per step ({cores} threads, contiguous chunks, one PRG per thread), SCL's verbatim "
              "shamirSecretShare + shamirRecoverP per secret")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_time / max(1, len(vals)), "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "reference" else "port",
                         "sample": sample, "build": flags,
                         "single_thread": {"value": one, "unit": UNIT, "cores": 1,
                                           "sample": f"{n_one} secrets; SCL is single-threaded: this is the reference as shipped"}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------- GPU arm helpers
class Bench:
    """State shared by the phases of the GPU arm (one process = one GPU)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.args = args
        self.torch, self.dist = torch, dist
        self.pkg = entry.load_package()
        self.sh = self.pkg.sharding
        self.B = self.pkg.binding
        self.rank, self.world, self.local_rank = self.sh.dist_env()
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            torch.cuda.set_device(self.local_rank)
            dist.init_process_group("nccl", rank=self.rank, world_size=self.world,
                                    device_id=torch.device("cuda", self.local_rank))
        self.dev = torch.device("cuda", self.local_rank)
        torch.cuda.set_device(self.dev)
        self.ctx = self.pkg.Context(self.local_rank)   # raises without a GPU / without libsclgpu.so
        self.ctx.use_torch_stream()
        # second and third contexts on side streams: the two lanes of the pipelined step
        self.sA, self.sB = torch.cuda.Stream(), torch.cuda.Stream()
        self.ctxA, self.ctxB = self.pkg.Context(self.local_rank), self.pkg.Context(self.local_rank)
        self.ctxA.set_stream(self.sA.cuda_stream)
        self.ctxB.set_stream(self.sB.cuda_stream)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x: float) -> float:
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def all_true(self, ok: bool) -> bool:
        t = self.torch.tensor([1 if ok else 0], dtype=self.torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return bool(t.item())

    def launches(self) -> int:
        return self.ctx.launch_count + self.ctxA.launch_count + self.ctxB.launch_count

    def close(self):
        for c in (self.ctxA, self.ctxB, self.ctx):
            c.close()


class Workload:
    """One rank's slice of a batch: secrets, two share-plane buffers, the reconstructed secrets."""

    def __init__(self, b: Bench, n_local: int, lo: int, planes=None):
        torch = b.torch
        self.b, self.N, self.lo = b, n_local, lo
        self.shard = b.sh.Shard(b.rank, b.world, lo, lo + n_local)
        self.first_block = b.sh.share_first_block(FIELD, T, 0, self.shard)
        self.sec_first = b.sh.random_first_block(FIELD, 0, self.shard)
        self.d_sec = torch.empty(n_local, dtype=torch.int64, device=b.dev)
        if planes is None:
            self.planes = [torch.empty((NPARTIES, n_local), dtype=torch.int64, device=b.dev) for _ in range(2)]
        else:  # views into buffers the caller already owns
            self.planes = [p.view(-1)[: NPARTIES * n_local].view(NPARTIES, n_local) for p in planes]
        self.d_out = torch.empty(n_local, dtype=torch.int64, device=b.dev)
        b.ctx.random_dev(FIELD, SEED_SECRETS, self.sec_first, n_local, self.d_sec)  # Vector::random(PRG("secrets")) slice
        # batch k of the run draws its coefficients where one PRG would be after k whole batches (all ranks): every
        # batch has fresh polynomials, so the two plane buffers never hold the same bytes
        self.batch_blocks = n_local * b.world * b.sh.blocks_per_share_call(FIELD, T)
        self.batch_in = [0, 0]   # which batch planes[i] holds

    def fb(self, k):
        return self.first_block + k * self.batch_blocks

    # --- the step, three ways
    def share(self, ctx, k):
        self.batch_in[k & 1] = k
        ctx.shamir_share_dev(FIELD, self.d_sec, self.N, T, NPARTIES, SEED_SHARE, self.fb(k), self.planes[k & 1], self.b.B.PARTY_MAJOR)

    def recover(self, ctx, k):
        ctx.recover_p_dev(FIELD, self.planes[k & 1], self.N, NPARTIES, self.d_out, self.b.B.PARTY_MAJOR)

    def step_sequential(self, k=0):
        self.share(self.b.ctx, k)
        self.recover(self.b.ctx, k)

    def step_fused(self, k=0):
        self.batch_in[k & 1] = k
        self.b.ctx.shamir_share_recover_dev(self.d_sec, self.N, T, NPARTIES, SEED_SHARE, self.fb(k), self.planes[k & 1], self.d_out)

    def step_fused_pipelined(self, k):
        self.batch_in[k & 1] = k
        self.b.ctx.shamir_share_recover_dev(self.d_sec, self.N, T, NPARTIES, SEED_SHARE, self.fb(k), self.planes[k & 1],
                                            self.d_out, rec_shares=self.planes[(k - 1) & 1])

    def time_one_stream(self, step, steps, split=False):
        """K steps on the current stream; -> (ms per step, share ms, recover ms) by CUDA events, this rank."""
        torch, b = self.b.torch, self.b
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
        b.barrier()
        ev[0].record()
        for k in range(steps):
            if split:
                self.share(b.ctx, k)
                ev[2 * k + 1].record()
                self.recover(b.ctx, k)
            else:
                step(k)
                ev[2 * k + 1].record()
            ev[2 * k + 2].record()
        b.barrier()
        total = ev[0].elapsed_time(ev[2 * steps]) / steps
        sh = sum(ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(steps)) / steps
        rc = sum(ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(steps)) / steps
        return total, sh, rc

    def time_two_streams(self, steps, warm):
        """Pipelined schedule: share(k) on stream A, recover(k-1) on stream B; -> ms per step (this rank)."""
        torch, b = self.b.torch, self.b
        sA, sB = b.sA, b.sB
        done_share = [None, None]   # event after the share that filled planes[i]
        done_rec = [None, None]     # event after the reconstruction that last read planes[i]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

        def one(k):
            i, j = k & 1, (k - 1) & 1
            if done_rec[i] is not None:
                sA.wait_event(done_rec[i])      # planes[i] was read by recover(k-2)
            with torch.cuda.stream(sA):
                self.share(b.ctxA, k)
                done_share[i] = torch.cuda.Event()
                done_share[i].record(sA)
            if done_share[j] is not None:
                sB.wait_event(done_share[j])    # recover(k-1) needs share(k-1)
                with torch.cuda.stream(sB):
                    self.recover(b.ctxB, k - 1)
                    done_rec[j] = torch.cuda.Event()
                    done_rec[j].record(sB)

        b.barrier()
        k = 0
        for _ in range(warm):
            one(k)
            k += 1
        sA.synchronize()
        sB.synchronize()
        b.barrier()
        l0 = b.launches()
        e0.record(sA)
        sB.wait_event(e0)
        for _ in range(steps):
            one(k)
            k += 1
        self.timed_launches = b.launches() - l0
        last = torch.cuda.Event()
        last.record(sB)
        sA.wait_event(last)
        e1.record(sA)
        sA.synchronize()
        sB.synchronize()
        ms = e0.elapsed_time(e1) / steps
        # drain: reconstruct the last batch too, so that d_out belongs to a complete round trip
        with torch.cuda.stream(sB):
            sB.wait_event(done_share[(k - 1) & 1])
            self.recover(b.ctxB, k - 1)
        sB.synchronize()
        b.barrier()
        return ms

    # --- verification against the oracle at this rank's own offsets
    def verify(self, port, which=(0, 1), K=1024) -> bool:
        import numpy as np

        torch = self.b.torch
        ok = bool(torch.equal(self.d_out, self.d_sec))
        spots = sorted({0, max(0, self.N // 2 - 5), max(0, self.N - K)})
        for lo in spots:
            k = min(K, self.N - lo)
            if k <= 0:
                continue
            lo -= lo % 2   # Vector::random draws two Fp61 elements per block
            sec_h = self.d_sec[lo:lo + k].cpu().numpy().view(np.uint64)
            ok = ok and bool(np.array_equal(sec_h, port.vector_random(FIELD, SEED_SECRETS, self.sec_first + lo // 2, k)))
            for w in which:
                want = port.shamir_share(FIELD, sec_h, T, NPARTIES, SEED_SHARE, self.fb(self.batch_in[w]) + lo * 8)
                got = self.planes[w][:, lo:lo + k].t().contiguous().cpu().numpy().view(np.uint64)
                ok = ok and bool(np.array_equal(got, want))
                ok = ok and bool(np.array_equal(self.d_out[lo:lo + k].cpu().numpy().view(np.uint64), port.recover_p(FIELD, want)))
        return ok


def check_gathered(b: Bench, ptr: int, n_total: int) -> bool:
    """This rank's copy of the gathered vector against the whole secret stream (regenerated on the device)."""
    torch = b.torch
    mine = torch.empty(n_total, dtype=torch.int64, device=b.dev)
    ref_sec = torch.empty(n_total, dtype=torch.int64, device=b.dev)
    b.ctx.random_dev(FIELD, SEED_SECRETS, 0, n_total, ref_sec)
    b.ctx.memcpy_d2d(mine.data_ptr(), ptr, 8 * n_total)
    torch.cuda.synchronize()
    same = bool(torch.equal(mine, ref_sec))
    # wipe the buffer so that the next variant's check cannot pass on stale data
    mine.zero_()
    b.ctx.memcpy_d2d(ptr, mine.data_ptr(), 8 * n_total)
    torch.cuda.synchronize()
    return same


def gathered_phase(b: Bench, wl: Workload, n_total: int, steps: int):
    """share + recoverP + all-gather of the reconstructed secrets, the gather inside the timing, three ways."""
    torch, dist = b.torch, b.dist
    res = {"bytes_gathered_per_rank": 8 * n_total, "secrets_total": n_total}
    # (a) NCCL all_gather after the reconstruction kernel, same stream
    gathered = torch.empty(n_total, dtype=torch.int64, device=b.dev)

    def step_nccl(k):
        wl.step_sequential(k)
        if b.world > 1:
            dist.all_gather_into_tensor(gathered, wl.d_out)
        else:
            gathered.copy_(wl.d_out)

    for k in range(2):
        step_nccl(k)
    ms, _, _ = wl.time_one_stream(step_nccl, steps)
    ms = b.max_over_ranks(ms)
    ok = bool(torch.equal(gathered[wl.lo:wl.lo + wl.N], wl.d_sec))
    res["nccl_all_gather"] = {"ms_per_step": ms, "value": n_total / (ms * 1e-3), "unit": UNIT,
                              "collective": "ncclAllGather after k_recover61_pm, same stream" if b.world > 1 else "device copy (1 rank)"}
    del gathered
    # (b) the gather fused into the reconstruction kernel: stores to every rank's buffer over NVLink peer memory
    ptr = b.ctx.malloc(8 * n_total)
    handles = [None] * b.world
    if b.world > 1:
        dist.all_gather_object(handles, b.ctx.ipc_export(ptr))
    peers = []
    for r in range(b.world):
        peers.append(ptr if r == b.rank else b.ctx.ipc_open(handles[r]))

    def step_p2p(k):
        wl.share(b.ctx, k)
        b.ctx.recover_p_gather_dev(wl.planes[k & 1], wl.N, NPARTIES, peers, wl.lo)

    for k in range(2):
        step_p2p(k)
    ms, _, _ = wl.time_one_stream(step_p2p, steps)
    ms = b.max_over_ranks(ms)
    b.barrier()
    ok = ok and check_gathered(b, ptr, n_total)
    b.barrier()

    # (c) ONE launch per step: share(batch k) + recoverP(batch k-1) + the gather, the NVLink stores running under the
    # share groups' work (k_share_recover61 with gather destinations)
    def step_one(k):
        wl.batch_in[k & 1] = k
        b.ctx.shamir_share_recover_gather_dev(wl.d_sec, wl.N, T, NPARTIES, SEED_SHARE, wl.fb(k), wl.planes[k & 1], peers, wl.lo,
                                              rec_shares=wl.planes[(k - 1) & 1])

    for k in range(2, 4):
        step_one(k)
    ms1 = b.max_over_ranks(wl.time_one_stream(lambda k: step_one(k + 4), steps)[0])
    b.barrier()
    res["one_launch_fused_p2p"] = {"ms_per_step": ms1, "value": n_total / (ms1 * 1e-3), "unit": UNIT,
                                   "collective": "none: k_share_recover61 shares batch k, reconstructs batch k-1 and stores each secret to "
                                                 "every rank's buffer (NVLink peer memory) in one persistent launch"}
    ok = ok and check_gathered(b, ptr, n_total)
    b.barrier()
    for r in range(b.world):
        if r != b.rank:
            b.ctx.ipc_close(peers[r])
    b.barrier()
    b.ctx.free(ptr)
    res["fused_p2p_stores"] = {"ms_per_step": ms, "value": n_total / (ms * 1e-3), "unit": UNIT,
                               "collective": "none: k_recover61_pm stores each secret to every rank's buffer (NVLink peer memory)"}
    res["verified_all_ranks"] = b.all_true(ok)
    return res


def run_b200(args):
    import numpy as np
    import torch

    b = Bench(args)
    rank, world = b.rank, b.world
    o = entry.load_oracle()
    port = o.PortOracle()          # seekable checker (pinned against the reference by tests/)
    steps, warm = args.steps, max(args.warmup, 3)
    n, t = NPARTIES, T
    NW = 1 << args.log2_secrets                      # weak: per GPU
    NS = (1 << args.log2_secrets) // world           # strong: per GPU
    wl = Workload(b, NW, rank * NW)

    for k in range(warm):
        wl.step_sequential(k)
    b.barrier()
    # integer-pipe peak at the clocks of this box (denominator of the int-mul roofline)
    imad_peak = b.ctx.pipe_microbench(0, 1 << 14)
    imadw_peak = b.ctx.pipe_microbench(1, 1 << 14)
    lds_peak = b.ctx.pipe_microbench(4, 1 << 14)      # conflict-free LDS.32 lane-operations per second

    sampler = ClockSampler(b.local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- two kernels back to back on one stream: per-kernel times for the roofline block
    seq_ms, share_ms, rec_ms = wl.time_one_stream(None, steps, split=True)
    seq_ms = b.max_over_ranks(seq_ms)
    ok_round = bool(torch.equal(wl.d_out, wl.d_sec))
    # ---- the timed region of `value`: one persistent launch per step, share(batch k) + recoverP(batch k-1)
    for k in range(warm):
        wl.step_fused_pipelined(k)
    l0 = b.launches()
    prim_ms = b.max_over_ranks(wl.time_one_stream(lambda k: wl.step_fused_pipelined(k + warm), steps)[0])
    launches = b.launches() - l0
    clocks = sampler.stop() if rank == 0 else None
    wl.recover(b.ctx, steps + warm - 1)   # drain: the batch shared last is reconstructed too (outside the timing)
    ok_prim = bool(torch.equal(wl.d_out, wl.d_sec)) and wl.verify(port)
    # ---- the same step on two streams (two kernels), and as one launch reconstructing the tiles it stores itself
    pipe_ms = b.max_over_ranks(wl.time_two_streams(steps, warm))
    ok_round = ok_round and bool(torch.equal(wl.d_out, wl.d_sec)) and wl.verify(port)
    for k in range(2):
        wl.step_fused(k)
    fused_ms = b.max_over_ranks(wl.time_one_stream(wl.step_fused, steps)[0])
    ok_fused = bool(torch.equal(wl.d_out, wl.d_sec)) and wl.verify(port)
    verified_all = b.all_true(ok_round and ok_prim and ok_fused)

    weak = {"ms_per_step": prim_ms, "value": world * NW / (prim_ms * 1e-3)}
    schedules = {
        "one_launch_pipelined": {"ms_per_step": prim_ms, "value": world * NW / (prim_ms * 1e-3), "launches_per_step": 1,
                                 "kernel": "k_share_recover61<4,4,tc>: four tcgen05 share groups + a reconstruction group on the tensor core; "
                                           "share(batch k) + recoverP(batch k-1); this is `value`"},
        "one_launch_same_batch": {"ms_per_step": fused_ms, "value": world * NW / (fused_ms * 1e-3), "launches_per_step": 1,
                                  "kernel": "k_share_recover61<4,4>: four share groups + four IMAD reconstruction warps, reconstruction of the tiles the launch itself stores"},
        "two_streams_pipelined": {"ms_per_step": pipe_ms, "value": world * NW / (pipe_ms * 1e-3), "launches_per_step": 2,
                                  "note": "k_share_tcm || k_recover61_pm: co-residency is up to the block scheduler"},
        "one_stream_back_to_back": {"ms_per_step": seq_ms, "value": world * NW / (seq_ms * 1e-3), "launches_per_step": 2,
                                    "share_ms": share_ms, "recover_ms": rec_ms},
        "verified": ok_round and ok_fused and ok_prim,
    }

    # ---- gathered (weak): the reconstructed secrets of all ranks on every rank, inside the timing
    gathered_weak = None
    if not args.no_gathered:
        gathered_weak = gathered_phase(b, wl, world * NW, max(3, steps // 2))

    # ---- strong scaling: 2^k secrets in total (BASELINE configs[1] as worded)
    if world == 1:
        strong = {"secrets_total": NW, "secrets_per_gpu": NW, "ms_per_step": prim_ms, "value": NW / (prim_ms * 1e-3),
                  "note": "one rank: identical to the weak line"}
        if gathered_weak is not None:
            strong["gathered"] = gathered_weak
        verified_strong = True
    else:
        ws = Workload(b, NS, rank * NS, planes=wl.planes)
        for k in range(warm):
            ws.step_sequential(k)
        s_seq, s_sh, s_rc = ws.time_one_stream(None, steps, split=True)
        s_seq = b.max_over_ranks(s_seq)
        for k in range(warm):
            ws.step_fused_pipelined(k)
        s_pipe = b.max_over_ranks(ws.time_one_stream(lambda k: ws.step_fused_pipelined(k + warm), steps)[0])
        ws.recover(b.ctx, steps + warm - 1)
        ok_s = bool(torch.equal(ws.d_out, ws.d_sec)) and ws.verify(port)
        strong = {"secrets_total": NS * world, "secrets_per_gpu": NS, "ms_per_step": s_pipe, "value": world * NS / (s_pipe * 1e-3),
                  "one_stream_back_to_back_ms": s_seq, "share_ms": s_sh, "recover_ms": s_rc}
        if not args.no_gathered:
            strong["gathered"] = gathered_phase(b, ws, world * NS, max(3, steps // 2))
            ok_s = ok_s and strong["gathered"]["verified_all_ranks"]
        verified_strong = b.all_true(ok_s)
        strong["verified_vs_oracle_all_ranks"] = verified_strong
        del ws

    # ---- staged form: coefficients already expanded in HBM by the PRG kernel, so the timed
    # region is Polynomial::evaluate at n points + shamirRecoverP.  Reported beside `value`, never as it.
    staged = None
    if not args.no_staged:
        N = NW
        d_planes = torch.empty((t + 1, N), dtype=torch.int64, device=b.dev)
        b.ctx.random_dev(FIELD, SEED_SHARE, wl.first_block, (t + 1) * N, d_planes)   # untimed PRG expansion
        d_planes[0].copy_(wl.d_sec)
        for _ in range(3):
            b.ctx.shamir_share_coeffs_dev(FIELD, d_planes, N, t, n, wl.planes[0], b.B.PARTY_MAJOR)
            wl.recover(b.ctx, 0)
        b.barrier()
        s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        sc_ms = sr_ms = 0.0
        for _ in range(steps):
            s0.record()
            b.ctx.shamir_share_coeffs_dev(FIELD, d_planes, N, t, n, wl.planes[0], b.B.PARTY_MAJOR)
            s1.record()
            wl.recover(b.ctx, 0)
            s2.record()
            torch.cuda.synchronize()
            sc_ms += s0.elapsed_time(s1) / steps
            sr_ms += s1.elapsed_time(s2) / steps
        b.barrier()
        ok_staged = bool(torch.equal(wl.d_out, wl.d_sec))
        st_ms = b.max_over_ranks(sc_ms + sr_ms)
        staged = {"value": world * N / (st_ms * 1e-3), "unit": UNIT, "ms_per_step": st_ms,
                  "share_from_coeffs_ms": sc_ms, "recover_ms": sr_ms, "verified": ok_staged,
                  "share_from_coeffs_GBps": (8 * (t + 1) + 8 * n) * N / (sc_ms * 1e-3) / 1e9,
                  "note": "coefficient planes pre-expanded in HBM (PRG outside the timed region); "
                          "k_share_tcm<F61,4,1,64,coeffs> (next tile prefetched) + k_recover61_pm<2>"}
        del d_planes
        torch.cuda.empty_cache()

    # ---- BASELINE's other configs (carved out of the two share-plane buffers)
    configs = None
    if not args.no_configs:
        configs = other_configs(b, port, o, wl.planes)
    del wl
    torch.cuda.empty_cache()

    # ---- e2e: host buffers through the reference-facing C ABI
    e2e = None
    if not args.no_e2e:
        e2e = e2e_phase(b, args, NW if args.scaling == "weak" else NS)

    if gathered_weak is not None:
        verified_all = verified_all and gathered_weak["verified_all_ranks"]
    verified_all = b.all_true(verified_all and verified_strong)
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        N = NW
        share_gbs = ALGO_BYTES_SHARE * N / (share_ms * 1e-3) / 1e9
        rec_gbs = ALGO_BYTES_RECOVER * N / (rec_ms * 1e-3) / 1e9
        share_kernel = {"0": "k_share61<15>", "1": "k_share61_tc", "2": "k_share_tcm<F61,4,1,64>"}.get(
            os.environ.get("SCLGPU_SHARE_TC", "3"), "k_share_tcm<F61,5,1,64>")
        # the step IS one kernel: its CUDA-event time is ms_per_step, its algorithmic bytes the whole 528 B per secret
        dominant = "k_share_recover61<4,4,tc>"
        dom_gbs = (ALGO_BYTES_SHARE + ALGO_BYTES_RECOVER) * N / (prim_ms * 1e-3) / 1e9
        traffic, traffic_src = None, None
        try:
            prof = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
            traffic = prof.get(dominant)
            traffic_src = prof.get("_source", "profiles/traffic.json")
        except (OSError, ValueError):
            pass
        prim = weak if args.scaling == "weak" else strong
        value, ms_per_step = prim["value"], prim["ms_per_step"]
        n_step = NW if args.scaling == "weak" else NS
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": launches, "verified_vs_oracle_all_ranks": verified_all,
            "verified_bit_exact_roundtrip": verified_all,
            "kernels": {"share_ms": share_ms, "recover_ms": rec_ms, "share_GBps": share_gbs, "recover_GBps": rec_gbs,
                        "note": "each kernel alone on one stream (CUDA events around every launch)"},
            "schedules": schedules,
            "weak": {"secrets_per_gpu": NW, "secrets_total": NW * world, **weak},
            "strong": strong,
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": dom_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": dom_gbs / hbm_peak, "traffic": traffic,
                         "traffic_source": f"profiled, static: {traffic_src} (ncu --set full capture of the same kernel; not measured in this run)",
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_secret": ALGO_BYTES_SHARE + ALGO_BYTES_RECOVER,
                         "share_of_step": 1.0,
                         "binding_resource": "NOT HBM: the SM's shared-memory data pipe (T-table AES-128-CTR lookups) and ALU pipe, "
                                             "both ~80-90 % busy (profiles/); `frac` is the HBM fraction the contract asks for",
                         "limiter": {"pipe": "lsu (shared-memory lookups of the fused AES-128-CTR)",
                                     "lookups_per_secret": AES_LDS_PER_SECRET,
                                     "achieved": AES_LDS_PER_SECRET * N / (prim_ms * 1e-3), "peak": lds_peak,
                                     "unit": "LDS.32 lane-ops/s", "frac": AES_LDS_PER_SECRET * N / (prim_ms * 1e-3) / lds_peak,
                                     "peak_source": "sclgpu_pipe_microbench(kind=4) on this GPU"},
                         "back_to_back_kernels": {
                             share_kernel: {"ms": share_ms, "achieved": share_gbs, "frac": share_gbs / hbm_peak,
                                            "algorithmic_bytes_per_secret": ALGO_BYTES_SHARE, "bound": "shared-memory + ALU pipes"},
                             "k_recover61_pm<2>": {"ms": rec_ms, "achieved": rec_gbs, "frac": rec_gbs / hbm_peak,
                                                   "algorithmic_bytes_per_secret": ALGO_BYTES_RECOVER, "bound": "hbm"}}},
            "int_roofline": {"unit": "IMAD/s", "algorithmic_imads_per_secret": ALGO_IMADS_PER_SECRET,
                             "achieved": ALGO_IMADS_PER_SECRET * n_step / (ms_per_step * 1e-3),
                             "peak_imad32": imad_peak, "peak_imad_wide": imadw_peak,
                             "frac": ALGO_IMADS_PER_SECRET * n_step / (ms_per_step * 1e-3) / imad_peak,
                             "frac_one_stream_back_to_back": ALGO_IMADS_PER_SECRET * NW / (seq_ms * 1e-3) / imad_peak,
                             "note": "per GPU; peak = measured by sclgpu_pipe_microbench on this GPU just before the timed region"},
        }
        if gathered_weak is not None:
            line["gathered"] = gathered_weak
        if staged is not None:
            staged["int_roofline_frac"] = ALGO_IMADS_PER_SECRET * NW / (staged["ms_per_step"] * 1e-3) / imad_peak
            staged["hbm_frac_share"] = staged["share_from_coeffs_GBps"] / hbm_peak
            line["staged"] = staged
        if configs is not None:
            for c in configs.values():
                if isinstance(c, dict) and "GBps" in c:
                    c["hbm_frac"] = c["GBps"] / hbm_peak
            line["configs"] = configs
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            v, kind, cores, n_sample, flags = cpu_reference_run(args.cpu_seconds)
            one, _, _, n_one, _ = cpu_reference_run(min(args.cpu_seconds, 4.0), threads=1)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "reference" else "port", "build": flags,
                "sample": f"{n_sample} secrets, SCL's per-secret shamirSecretShare+shamirRecoverP on {cores} threads "
                          f"(about {args.cpu_seconds:.0f} s of CPU work)",
                "single_thread": {"value": one, "unit": UNIT, "cores": 1,
                                  "sample": f"{n_one} secrets; SCL is single-threaded: this is the reference as shipped"}}
        emit(line)
    b.close()
    if world > 1:
        b.dist.barrier()
        b.dist.destroy_process_group()


# ----------------------------------------------------------------- other BASELINE configs
def other_configs(b: Bench, port, o, scratch):
    """C1, C3, C4, C5 on the device, each with its own algorithmic bytes and an oracle check.  `scratch`: two big int64
    buffers to carve from (no new allocation of that size).  At world > 1 only C5 (row-sharded) uses the other ranks."""
    import numpy as np

    torch, ctx, B = b.torch, b.ctx, b.B
    res = {}
    reps = 5

    def timeit(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    flat = [s.view(-1) for s in scratch]
    need = 1 << 29
    for i in range(2):
        if flat[i].numel() < need:   # reduced --log2-secrets: the plane buffers are too small to carve from
            flat[i] = torch.empty(need, dtype=torch.int64, device=b.dev)

    def carve(buf, off, *shape):
        cnt = int(np.prod(shape))
        return flat[buf][off:off + cnt].view(*shape), off + cnt

    try:
        orc = o.best_oracle()
    except Exception:
        orc = port
    if b.rank == 0:
        # ---- C1: Fp61 n=5 t=2, 2^20 secrets (the configuration SCL runs on a CPU today): full check vs the oracle
        N, n, t = 1 << 20, 5, 2
        sec, off = carve(0, 0, N)
        sh, off = carve(0, off, n, N)
        out, off = carve(0, off, N)
        ctx.random_dev(61, SEED_SECRETS, 0, N, sec)
        ms_sh = timeit(lambda: ctx.shamir_share_dev(61, sec, N, t, n, SEED_SHARE, 0, sh, B.PARTY_MAJOR))
        ms_rc = timeit(lambda: ctx.recover_p_dev(61, sh, N, n, out, B.PARTY_MAJOR))
        ms_one = timeit(lambda: ctx.shamir_share_recover_dev(sec, N, t, n, SEED_SHARE, 0, sh, out))
        sec_h = sec.cpu().numpy().view(np.uint64)
        want = orc.shamir_share(61, sec_h, t, n, SEED_SHARE, 0)
        ok = bool(np.array_equal(sh.t().contiguous().cpu().numpy().view(np.uint64), want)) and bool(torch.equal(out, sec)) \
            and bool(np.array_equal(out.cpu().numpy().view(np.uint64), orc.recover_p(61, want)))
        best = min(ms_sh + ms_rc, ms_one)
        res["C1_fp61_n5_t2_2^20"] = {
            "share_ms": ms_sh, "recover_ms": ms_rc, "one_launch_ms": ms_one, "value": N / (best * 1e-3), "unit": UNIT,
            "algorithmic_bytes_per_secret": 96, "GBps": 96 * N / (best * 1e-3) / 1e9,
            "int_imads_per_secret": 60, "note": "48 MB working set: L2-resident, launch-latency bound",
            "verified_vs_oracle": ok, "checked": f"all {N} sharings vs oracle ({orc.kind})"}

        # ---- C3: Fp127 n=16 t=7, 2^24 secrets, error-detecting reconstruction + tamper set
        N, n, t = 1 << 24, 16, 7
        sec, off = carve(0, 0, N, 2)
        out, off = carve(0, off, N, 2)
        sh, _ = carve(1, 0, n, N, 2)
        err = torch.empty(N, dtype=torch.uint8, device=b.dev)
        ctx.random_dev(127, "secrets127", 0, N, sec)
        ms_sh = timeit(lambda: ctx.shamir_share_dev(127, sec, N, t, n, SEED_SHARE, 0, sh, B.PARTY_MAJOR))
        ms_rd = timeit(lambda: ctx.recover_d_dev(127, sh, N, n, t, out, err, B.PARTY_MAJOR))
        ok = bool(torch.equal(out, sec)) and int(err.sum().item()) == 0
        K = 2048
        sec_h = sec[:K].cpu().numpy().view(np.uint64)
        want = orc.shamir_share(127, sec_h, t, n, SEED_SHARE, 0)
        ok = ok and bool(np.array_equal(sh[:, :K].permute(1, 0, 2).contiguous().cpu().numpy().view(np.uint64), want))
        # tamper set: flip one share at idx in {0,8,13,14,15} for 1/1024 of the secrets
        idxs = [0, 8, 13, 14, 15]
        for q, idx in enumerate(idxs):
            sh[idx, q::5 * 1024, 0] ^= 1
        nd = ctx.recover_d_dev(127, sh, N, n, t, out, err, B.PARTY_MAJOR)
        got_h = sh[:, :8 * 5 * 1024].permute(1, 0, 2).contiguous().cpu().numpy().view(np.uint64)
        w_out, w_err, w_nd = orc.recover_d(127, got_h, t)
        ok = ok and bool(np.array_equal(err[:8 * 5 * 1024].cpu().numpy(), w_err)) \
            and bool(np.array_equal(out[:8 * 5 * 1024].cpu().numpy().view(np.uint64), w_out))
        flagged = int(err.sum().item())
        ok = ok and flagged == nd
        res["C3_fp127_n16_t7_2^24_recoverD"] = {
            "share_ms": ms_sh, "recover_d_ms": ms_rd, "value": N / ((ms_sh + ms_rd) * 1e-3), "unit": UNIT,
            "algorithmic_bytes_per_secret": 512, "GBps": 512 * N / ((ms_sh + ms_rd) * 1e-3) / 1e9,
            "int_imads_per_secret": 2688, "tampered": sum(len(range(q, N, 5 * 1024)) for q in range(5)), "flagged": flagged,
            "note": "flags: idx 0/8/13 detected, idx 14/15 not (the reference's own loop bound, shamir.h:129)",
            "verified_vs_oracle": ok, "checked": f"prefix {K} shares + first 40960 tampered secrets/flags vs oracle ({orc.kind})"}
        del err

        # ---- C4: AES-CTR keystream -> 2^28 Fp61 elements (2 GiB)
        n4 = 1 << 28
        buf, _ = carve(0, 0, n4)
        ms_ks = timeit(lambda: ctx.prg_expand_dev("prg bench", 0, 8 * n4, buf))
        ks_head = buf[:4096].cpu().numpy().view(np.uint8).copy()
        far = (n4 - 4096)
        ks_far = buf[far:].cpu().numpy().view(np.uint8).copy()
        # the same keystream from the bitsliced kernel (no table lookups): the measured comparison arm
        ms_bs = timeit(lambda: ctx.prg_expand_bitsliced_dev("prg bench", 0, 8 * n4, buf))
        bs_ok = bool(np.array_equal(buf[:4096].cpu().numpy().view(np.uint8), ks_head)) \
            and bool(np.array_equal(buf[far:].cpu().numpy().view(np.uint8), ks_far))
        ms_r = timeit(lambda: ctx.random_dev(61, "prg bench", 0, n4, buf))
        ok = bs_ok and bool(np.array_equal(ks_head, port.prg_next("prg bench", 0, 8 * 4096))) \
            and bool(np.array_equal(ks_far, port.prg_next("prg bench", far // 2, 8 * 4096))) \
            and bool(np.array_equal(buf[:4096].cpu().numpy().view(np.uint64), orc.vector_random(61, "prg bench", 0, 4096))) \
            and bool(np.array_equal(buf[far:].cpu().numpy().view(np.uint64), port.vector_random(61, "prg bench", far // 2, 4096)))
        res["C4_prg_2^28_fp61"] = {
            "keystream_ms": ms_ks, "fp61_random_ms": ms_r, "value": (n4 / 2) / (ms_ks * 1e-3), "unit": "AES blocks/s",
            "GBps": 8 * n4 / (ms_ks * 1e-3) / 1e9, "elements_per_s": n4 / (ms_r * 1e-3),
            "bound": "shared-memory data pipe + ALU pipe (T-table AES), not HBM",
            "lds_frac": 133.0 * (n4 / 2) / (ms_ks * 1e-3) / b.ctx.pipe_microbench(4, 1 << 14),
            "bitsliced": {"keystream_ms": ms_bs, "value": (n4 / 2) / (ms_bs * 1e-3), "unit": "AES blocks/s",
                          "kernel": "k_prg_bitsliced: Boyar-Peralta S-box circuit, 32 blocks per thread, no lookups",
                          "vs_t_table": ms_bs / ms_ks, "verified": bs_ok,
                          "lop3_per_s_measured": b.ctx.pipe_microbench(2, 1 << 14)},
            "verified_vs_oracle": ok, "checked": "first and last 32 KiB of keystream and of Vector::random vs oracle"}

    # ---- C5: Fp61 mat-vec 8192 x 8192 (rows sharded over the ranks) and Beaver mul-add on 2^26 elements
    rows = cols = 8192
    shard = b.sh.shard_range(rows, b.world, b.rank)
    A, off = carve(0, 0, shard.count, cols)
    x, off = carve(0, off, cols)
    ctx.random_dev(61, "mat A", shard.lo * cols // 2, shard.count * cols, A)
    ctx.random_dev(61, "vec x", 0, cols, x)
    y_local = torch.empty(shard.count, dtype=torch.int64, device=b.dev)
    y = torch.empty(rows, dtype=torch.int64, device=b.dev)

    def mv():
        ctx.matvec_dev(61, A, shard.count, cols, x, y_local)
        if b.world > 1:
            b.dist.all_gather_into_tensor(y, y_local)

    if b.world == 1:
        y_local = y
    ms_mv = b.max_over_ranks(timeit(mv))
    ok5 = True
    if b.rank == 0:
        A_h = A[:64].cpu().numpy().view(np.uint64)
        ok5 = bool(np.array_equal(y[:64].cpu().numpy().view(np.uint64), orc.matvec(61, A_h, x.cpu().numpy().view(np.uint64))))
    r5 = {"rows": rows, "cols": cols, "rows_per_gpu": shard.count, "matvec_ms": ms_mv,
          "value": rows * cols / (ms_mv * 1e-3), "unit": "modmul/s", "GBps": 8 * rows * cols / (ms_mv * 1e-3) / 1e9,
          "collective": "ncclAllGather of the y slices (8 KiB per GPU at 8 ranks), inside the timing" if b.world > 1 else "none"}
    if b.rank == 0:
        n5 = 1 << 26
        vs = []
        off = 0
        for k in range(6):
            v, off = carve(1, off, n5)
            vs.append(v)
        for k, v in enumerate(vs[:5]):
            ctx.random_dev(61, "ebdac"[k], 0, n5, v)
        ms_ma = timeit(lambda: ctx.beaver_dev(61, vs[0], vs[1], vs[2], vs[3], vs[4], n5, vs[5]))
        K = 4096
        h = [v[:K].cpu().numpy().view(np.uint64) for v in vs[:5]]
        ok5 = ok5 and bool(np.array_equal(vs[5][:K].cpu().numpy().view(np.uint64), orc.beaver(61, *h)))
        r5.update({"muladd_elements": n5, "muladd_ms": ms_ma, "muladd_GBps": 48 * n5 / (ms_ma * 1e-3) / 1e9,
                   "muladd_elements_per_s": n5 / (ms_ma * 1e-3)})
    r5["verified_vs_oracle"] = b.all_true(ok5)
    r5["checked"] = "first 64 rows of y and first 4096 mul-add elements vs oracle"
    res["C5_fp61_matvec_8192_muladd_2^26"] = r5
    return res


# ----------------------------------------------------------------- e2e
def pcie_probe(b: Bench, nbytes=1 << 30, reps=3):
    """Concurrent pinned H2D + D2H on this rank's GPU (all ranks at the same time): the box's ceiling for the e2e
    path under the same concurrency.  -> (h2d GB/s, d2h GB/s) of this rank."""
    torch = b.torch
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=b.dev)
    d_out = torch.empty(nbytes, dtype=torch.uint8, device=b.dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    b.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)
    s1.synchronize()
    s2.synchronize()
    dt = time.perf_counter() - t0
    b.barrier()
    return reps * nbytes / dt / 1e9, reps * nbytes / dt / 1e9


def e2e_phase(b: Bench, args, n_dev: int):
    import ctypes as C

    import numpy as np

    torch, ctx, pkg = b.torch, b.ctx, b.pkg
    n, t = NPARTIES, T
    # host footprint guard: every rank pins 8*N*(2n+2) bytes (two share buffers); keep the sum below half of MemAvailable
    lg_e = n_dev.bit_length() - 1
    try:
        avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
    except (OSError, StopIteration):
        avail = 1 << 40
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", b.world))
    while lg_e > 16 and local_world * 8 * (1 << lg_e) * (2 * n + 2) > avail // 2:
        lg_e -= 1
    N = 1 << lg_e
    lo = b.rank * N
    shard = b.sh.Shard(b.rank, b.world, lo, lo + N)
    first_block = b.sh.share_first_block(FIELD, T, 0, shard)
    h_sec = ctx.host_alloc(8 * N).view(np.uint64)
    h_sh = [ctx.host_alloc(8 * N * n).view(np.uint64) for _ in range(2)]
    h_out = ctx.host_alloc(8 * N).view(np.uint64)
    d_sec = torch.empty(N, dtype=torch.int64, device=b.dev)
    ctx.random_dev(FIELD, SEED_SECRETS, b.sh.random_first_block(FIELD, 0, shard), N, d_sec)
    h_sec[:] = d_sec.cpu().numpy().view(np.uint64)
    del d_sec

    def p(a):
        return a.ctypes.data_as(C.c_void_p)

    seed = pkg.api.seed16(SEED_SHARE)
    lib = ctx.lib

    # (1) one direction at a time (round-1 definition): share, then reconstruct, synchronous calls
    def seq_step():
        ctx._check(lib.sclgpu_fp61_shamir_share(ctx._ctx, p(h_sec), N, t, n, seed, first_block, p(h_sh[0])))
        ctx._check(lib.sclgpu_fp61_recover_p(ctx._ctx, p(h_sh[0]), N, n, None, None, p(h_out)))

    seq_step()  # warm-up (allocations, basis cache)
    b.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        seq_step()
    torch.cuda.synchronize()
    seq_s = b.max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    ok = bool(np.array_equal(h_out, h_sec))

    # (2) full duplex: share_async(batch k) while recover_p(batch k-1); one step = one share + one reconstruction
    def dup_step(k):
        ctx._check(lib.sclgpu_fp61_shamir_share_async(ctx._ctx, p(h_sec), N, t, n, seed, first_block, p(h_sh[k & 1])))
        ctx._check(lib.sclgpu_fp61_recover_p(ctx._ctx, p(h_sh[(k - 1) & 1]), N, n, None, None, p(h_out)))
        ctx._check(lib.sclgpu_wait(ctx._ctx))

    ctx._check(lib.sclgpu_fp61_shamir_share(ctx._ctx, p(h_sec), N, t, n, seed, first_block, p(h_sh[1])))
    h_out[:] = 0
    dup_step(0)   # warm-up: companion context, its scratch
    b.barrier()
    t0 = time.perf_counter()
    for k in range(1, 1 + args.e2e_steps):
        dup_step(k)
    torch.cuda.synchronize()
    dup_s = b.max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    ok = ok and bool(np.array_equal(h_out, h_sec))
    # the buffer filled last holds a complete sharing: reconstruct it as the final check
    ctx._check(lib.sclgpu_fp61_recover_p(ctx._ctx, p(h_sh[args.e2e_steps & 1]), N, n, None, None, p(h_out)))
    ok = ok and bool(np.array_equal(h_out, h_sec))

    bytes_h2d = 8 * N + 8 * N * n
    bytes_d2h = 8 * N * n + 8 * N
    up, down = pcie_probe(b)
    ceil_up, ceil_down = b.sum_over_ranks(up), b.sum_over_ranks(down)
    e2e = {"value": b.world * N / dup_s, "unit": UNIT, "h2d_bytes_per_step": bytes_h2d, "d2h_bytes_per_step": bytes_d2h,
           "ms_per_step": 1e3 * dup_s, "steps": args.e2e_steps, "secrets_per_gpu": N, "verified": b.all_true(ok),
           "api": "sclgpu_fp61_shamir_share_async(batch k) || sclgpu_fp61_recover_p(batch k-1) + sclgpu_wait; pinned host "
                  "buffers, SCL [N][n] layout; one share and one reconstruction per step",
           "one_direction_at_a_time": {"value": b.world * N / seq_s, "ms_per_step": 1e3 * seq_s,
                                        "api": "sclgpu_fp61_shamir_share then sclgpu_fp61_recover_p (synchronous)"},
           "host_ceiling": {"h2d_GBps": ceil_up, "d2h_GBps": ceil_down,
                            "how": "all ranks at once: pinned cudaMemcpyAsync both directions, 1 GiB x 3"},
           "achieved_GBps_each_way": b.world * bytes_h2d / dup_s / 1e9}
    e2e["frac_of_host_ceiling"] = e2e["achieved_GBps_each_way"] / max(1e-9, min(ceil_up, ceil_down))
    for a in h_sh:
        ctx.host_free(a.view(np.uint8))

    # (3) the per-party packet entry points (Serializer<Vector<FF>> wire layout): no [N][n] matrix, no transposition
    pk_bytes = int(lib.sclgpu_packet_bytes(8, N))
    h_pk = [ctx.host_alloc(pk_bytes) for _ in range(n)]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in h_pk])

    def pk_step():
        ctx._check(lib.sclgpu_fp61_shamir_share_packets(ctx._ctx, p(h_sec), N, t, n, seed, first_block, ptrs))
        ctx._check(lib.sclgpu_fp61_recover_p_packets(ctx._ctx, ptrs, N, n, None, None, p(h_out)))

    h_out[:] = 0
    pk_step()
    b.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        pk_step()
    torch.cuda.synchronize()
    pk_s = b.max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
    e2e["packets"] = {"value": b.world * N / pk_s, "unit": UNIT, "ms_per_step": 1e3 * pk_s,
                      "verified": b.all_true(bool(np.array_equal(h_out, h_sec))),
                      "api": "sclgpu_fp61_shamir_share_packets + sclgpu_fp61_recover_p_packets (n pinned packet buffers)"}
    for a in h_pk:
        ctx.host_free(a)

    # (4) ONE process, ONE call over all the GPUs of the box (sclgpu_multi_*): rank 0 drives, the others wait
    if b.world > 1:
        b.barrier()
        sp = None
        if b.rank == 0:
            Nt = b.world * N
            h_all_sec = ctx.host_alloc(8 * Nt).view(np.uint64)
            h_all_sh = ctx.host_alloc(8 * Nt * n).view(np.uint64)
            h_all_out = ctx.host_alloc(8 * Nt).view(np.uint64)
            m = pkg.MultiContext(list(range(b.world)))
            try:
                h_all_sec[:N] = h_sec
                h_all_sec[N:] = m.random(SEED_SECRETS, N // 2, Nt - N)
                m.shamir_share(61, h_all_sec, t, n, SEED_SHARE, 0, out=h_all_sh.reshape(Nt, n))   # warm-up
                t0 = time.perf_counter()
                for _ in range(args.e2e_steps):
                    m.shamir_share(61, h_all_sec, t, n, SEED_SHARE, 0, out=h_all_sh.reshape(Nt, n))
                    m.recover_p(61, h_all_sh.reshape(Nt, n), out=h_all_out)
                sp_s = (time.perf_counter() - t0) / args.e2e_steps
                sp = {"value": Nt / sp_s, "unit": UNIT, "ms_per_step": 1e3 * sp_s, "secrets_total": Nt, "devices": b.world,
                      "verified": bool(np.array_equal(h_all_out, h_all_sec)),
                      "api": "sclgpu_multi_fp61_shamir_share + sclgpu_multi_fp61_recover_p: one process, one call, "
                             "contiguous slices over the devices, one worker thread per device"}
            finally:
                m.close()
            for a in (h_all_sec, h_all_sh, h_all_out):
                ctx.host_free(a.view(np.uint8))
        b.barrier()
        e2e["single_process_all_gpus"] = sp
    for a in (h_sec, h_out):
        ctx.host_free(a.view(np.uint8))
    return e2e


_REAL_STDOUT = None


def emit(line: dict):
    """The contract is ONE JSON line on stdout: libraries that write to fd 1 on their own (NCCL prints its version
    banner there) are kept off it by pointing fd 1 at stderr for the whole run and printing the line on the saved fd."""
    sys.stdout.flush()
    text = json.dumps(line) + "\n"
    if _REAL_STDOUT is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, text.encode())


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
