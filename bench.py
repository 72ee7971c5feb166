#!/usr/bin/env python
"""bench.py -- Fp61 Shamir share + reconstruct throughput (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic secrets:
    shares  = shamirSecretShare(secret_j, t, n, prg)   for all j   (PRG fused in)
    secret' = shamirRecoverP(shares_j)                 for all j
Workload at N=1: BASELINE configs[1] -- Mersenne-61, n=32, t=15, 2^26 secrets.
Multi-GPU: the batch-of-secrets dimension is sharded, one process per GPU, no
data-path collective; weak scaling: every rank works on 2^26 secrets,
rank r taking slice r of the global batch with the PRG counter offset to match.

  value  : secrets/s, inputs (secrets) resident in HBM, shares written to and
           read back from HBM (party-major), CUDA events, max over ranks.
  e2e    : the same metric through the reference-facing C-ABI host entry points
           sclgpu_fp61_shamir_share + sclgpu_fp61_recover_p with HOST (pinned)
           buffers in SCL's own [N][n] layout: H2D/D2H copies inside the timing.
  --impl reference : SCL's own CPU code (oracle/_ref, built from the unmodified
           reference sources) on all host cores, a bounded sample per step.
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
AES_LDS_PER_SECRET = 1083          # 8 blocks x 133 T-table lookups + 27 per-group lookups / ... (ncu op mix, profiles/)
ALGO_BYTES_SHARE = 8 + 8 * NPARTIES    # secret in, n shares out
ALGO_BYTES_RECOVER = 8 * NPARTIES + 8  # n shares in, secret out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log2-secrets", type=int, default=26, help="secrets per GPU = 2^k (default 26 = BASELINE configs[1])")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-staged", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def workload_config(args, world):
    return {
        "workload": f"Mersenne-61 Shamir share+recoverP n={NPARTIES} t={T}, 2^{args.log2_secrets} secrets per GPU"
                    " (BASELINE configs[1])",
        "field": "Fp61", "n": NPARTIES, "t": T, "secrets_per_gpu": 1 << args.log2_secrets,
        "secrets_total": (1 << args.log2_secrets) * world, "sharding": f"batch x{world}, no collective",
        "prg": "AES-128-CTR fused into the share kernel (seed 'shamir bench'): coefficients are drawn inside the timed region",
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
    Returns (secrets_per_s, kind, cores, n_sample)."""
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
    return n_sample / dt, orc.kind, cores, n_sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = max(1.0, min(10.0, 150.0 / max(1, args.steps + args.warmup)))
    vals = []
    kind, cores, n_sample = "port", 1, 0
    for i in range(args.warmup + args.steps):
        v, kind, cores, n_sample = cpu_reference_run(per_step)
        if i >= args.warmup:
            vals.append((v, n_sample))
    tot_secrets = sum(n for _, n in vals)
    tot_time = sum(n / v for v, n in vals)
    value = tot_secrets / tot_time
    sample = (f"{n_sample} secrets per step ({cores} threads, contiguous chunks, one PRG per thread), SCL's verbatim "
              "shamirSecretShare + shamirRecoverP per secret")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_time / max(1, len(vals)), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "reference" else "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    pkg = entry.load_package()
    sh = pkg.sharding
    B = pkg.binding
    rank, world, local_rank = sh.dist_env()
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    ctx = pkg.Context(local_rank)   # raises without a GPU / without libsclgpu.so
    ctx.use_torch_stream()

    N = 1 << args.log2_secrets
    n, t = NPARTIES, T
    shard = sh.Shard(rank, world, rank * N, (rank + 1) * N)   # weak scaling: slice r of a world*N batch
    first_block = sh.share_first_block(FIELD, t, 0, shard)
    sec_first = sh.random_first_block(FIELD, 0, shard)

    d_sec = torch.empty(N, dtype=torch.int64, device=dev)
    d_sh = torch.empty((n, N), dtype=torch.int64, device=dev)   # party-major share planes
    d_out = torch.empty(N, dtype=torch.int64, device=dev)
    ctx.random_dev(FIELD, "secrets", sec_first, N, d_sec)        # synthetic secrets = Vector::random(PRG("secrets"))

    def step():
        ctx.shamir_share_dev(FIELD, d_sec, N, t, n, "shamir bench", first_block, d_sh, B.PARTY_MAJOR)
        ctx.recover_p_dev(FIELD, d_sh, N, n, d_out, B.PARTY_MAJOR)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # integer-pipe peak at the clocks of this box (denominator of the int-mul roofline)
    imad_peak = ctx.pipe_microbench(0, 1 << 14)
    imadw_peak = ctx.pipe_microbench(1, 1 << 14)
    lds_peak = ctx.pipe_microbench(4, 1 << 14)      # conflict-free LDS.32 lane-operations per second

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * args.steps + 1)]
    launches0 = ctx.launch_count
    barrier()
    ev[0].record()
    for k in range(args.steps):
        ctx.shamir_share_dev(FIELD, d_sec, N, t, n, "shamir bench", first_block, d_sh, B.PARTY_MAJOR)
        ev[3 * k + 1].record()
        ctx.recover_p_dev(FIELD, d_sh, N, n, d_out, B.PARTY_MAJOR)
        ev[3 * k + 2].record()
        ev[3 * k + 3].record()
    barrier()
    launches = ctx.launch_count - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[3 * args.steps])
    share_ms = sum(ev[3 * k].elapsed_time(ev[3 * k + 1]) for k in range(args.steps)) / args.steps
    rec_ms = sum(ev[3 * k + 1].elapsed_time(ev[3 * k + 2]) for k in range(args.steps)) / args.steps
    verified = bool(torch.equal(d_out, d_sec))

    tms = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    total_ms = float(tms.item())
    ms_per_step = total_ms / args.steps
    value = world * N / (ms_per_step * 1e-3)

    # ---- staged form: coefficients already expanded in HBM by the PRG kernel, so the timed
    # region is Polynomial::evaluate at n points + shamirRecoverP.  Reported beside `value`, never as it.
    staged = None
    if not args.no_staged:
        d_planes = torch.empty((t + 1, N), dtype=torch.int64, device=dev)
        ctx.random_dev(FIELD, "shamir bench", first_block, (t + 1) * N, d_planes)   # untimed PRG expansion
        d_planes[0].copy_(d_sec)
        for _ in range(3):
            ctx.shamir_share_coeffs_dev(FIELD, d_planes, N, t, n, d_sh, B.PARTY_MAJOR)
            ctx.recover_p_dev(FIELD, d_sh, N, n, d_out, B.PARTY_MAJOR)
        barrier()
        s0, s1, s2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        sc_ms = sr_ms = 0.0
        for _ in range(args.steps):
            s0.record()
            ctx.shamir_share_coeffs_dev(FIELD, d_planes, N, t, n, d_sh, B.PARTY_MAJOR)
            s1.record()
            ctx.recover_p_dev(FIELD, d_sh, N, n, d_out, B.PARTY_MAJOR)
            s2.record()
            torch.cuda.synchronize()
            sc_ms += s0.elapsed_time(s1) / args.steps
            sr_ms += s1.elapsed_time(s2) / args.steps
        barrier()
        verified = verified and bool(torch.equal(d_out, d_sec))
        tst = torch.tensor([sc_ms + sr_ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tst, op=dist.ReduceOp.MAX)
        st_ms = float(tst.item())
        staged = {"value": world * N / (st_ms * 1e-3), "unit": UNIT, "ms_per_step": st_ms,
                  "share_from_coeffs_ms": sc_ms, "recover_ms": sr_ms,
                  "share_from_coeffs_GBps": (8 * (t + 1) + 8 * n) * N / (sc_ms * 1e-3) / 1e9,
                  "note": "coefficient planes pre-expanded in HBM (PRG outside the timed region); "
                          "k_share_tcm<F61,4,1,64,coeffs> (next tile prefetched) + k_recover61_pm<2>"}
        del d_planes
        torch.cuda.empty_cache()

    # ---- e2e: host buffers through the reference-facing C ABI
    e2e = None
    if not args.no_e2e:
        # host footprint guard: every rank pins 8*N*(n+2) bytes; keep the sum below half of MemAvailable
        N_full, lg_e = N, args.log2_secrets
        try:
            avail = next(int(l.split()[1]) * 1024 for l in open("/proc/meminfo") if l.startswith("MemAvailable"))
        except (OSError, StopIteration):
            avail = 1 << 40
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", world))
        while lg_e > 16 and local_world * 8 * (1 << lg_e) * (n + 2) > avail // 2:
            lg_e -= 1
        N = 1 << lg_e
        h_sec = ctx.host_alloc(8 * N).view(np.uint64)
        h_sh = ctx.host_alloc(8 * N * n).view(np.uint64)
        h_out = ctx.host_alloc(8 * N).view(np.uint64)
        h_sec[:] = d_sec[:N].cpu().numpy().view(np.uint64)
        import ctypes as C

        def p(a):
            return a.ctypes.data_as(C.c_void_p)

        seed = pkg.api.seed16("shamir bench")

        def e2e_step():
            rc = ctx.lib.sclgpu_fp61_shamir_share(ctx._ctx, p(h_sec), N, t, n, seed, first_block, p(h_sh))
            ctx._check(rc)
            rc = ctx.lib.sclgpu_fp61_recover_p(ctx._ctx, p(h_sh), N, n, None, None, p(h_out))
            ctx._check(rc)

        e2e_step()  # warm-up (allocations, basis cache)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        e_s = (time.perf_counter() - t0) / args.e2e_steps
        te = torch.tensor([e_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_s = float(te.item())
        verified = verified and bool(np.array_equal(h_out, h_sec))
        e2e = {"value": world * N / e_s, "unit": UNIT, "h2d_bytes_per_step": 8 * N + 8 * N * n,
               "d2h_bytes_per_step": 8 * N * n + 8 * N, "ms_per_step": 1e3 * e_s, "steps": args.e2e_steps,
               "secrets_per_gpu": N,
               "api": "sclgpu_fp61_shamir_share + sclgpu_fp61_recover_p, pinned host buffers, SCL [N][n] layout"}
        for a in (h_sh,):
            ctx.host_free(a.view(np.uint8))
        # the same round trip through the per-party packet entry points (Serializer<Vector<FF>> wire
        # layout, what a dealer sends / a reconstructing party receives): no [N][n] matrix, no transposition
        pk_bytes = int(ctx.lib.sclgpu_packet_bytes(8, N))
        h_pk = [ctx.host_alloc(pk_bytes) for _ in range(n)]
        ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in h_pk])

        def pk_step():
            ctx._check(ctx.lib.sclgpu_fp61_shamir_share_packets(ctx._ctx, p(h_sec), N, t, n, seed, first_block, ptrs))
            ctx._check(ctx.lib.sclgpu_fp61_recover_p_packets(ctx._ctx, ptrs, N, n, None, None, p(h_out)))

        h_out[:] = 0
        pk_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            pk_step()
        torch.cuda.synchronize()
        pk_s = (time.perf_counter() - t0) / args.e2e_steps
        tp = torch.tensor([pk_s], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        pk_s = float(tp.item())
        verified = verified and bool(np.array_equal(h_out, h_sec))
        e2e["packets"] = {"value": world * N / pk_s, "unit": UNIT, "ms_per_step": 1e3 * pk_s,
                          "api": "sclgpu_fp61_shamir_share_packets + sclgpu_fp61_recover_p_packets (n pinned packet buffers)"}
        for a in [h_sec, h_out] + h_pk:
            ctx.host_free(a.view(np.uint8))
        N = N_full

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json")))
        except OSError:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
        share_gbs = ALGO_BYTES_SHARE * N / (share_ms * 1e-3) / 1e9
        rec_gbs = ALGO_BYTES_RECOVER * N / (rec_ms * 1e-3) / 1e9
        share_kernel = {"0": "k_share61<15>", "1": "k_share61_tc", "2": "k_share_tcm<F61,4,1,64>"}.get(
            os.environ.get("SCLGPU_SHARE_TC", "3"), "k_share_tcm<F61,5,1,64>")
        dominant = share_kernel if share_ms >= rec_ms else "k_recover61_pm<2>"
        dom_gbs = share_gbs if share_ms >= rec_ms else rec_gbs
        traffic = None
        try:
            prof = json.load(open(os.path.join(REPO, "profiles", "traffic.json")))
            traffic = prof.get(dominant)
        except (OSError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u64", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks, "gpu_launches": launches, "verified_bit_exact_roundtrip": verified,
            "kernels": {"share_ms": share_ms, "recover_ms": rec_ms, "share_GBps": share_gbs, "recover_GBps": rec_gbs},
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": dom_gbs, "peak": hbm_peak, "unit": "GB/s",
                         "frac": dom_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_secret": ALGO_BYTES_SHARE if share_ms >= rec_ms else ALGO_BYTES_RECOVER,
                         "note": "the share kernel's limiter is the shared-memory (LSU) pipe of the fused T-table AES-CTR, "
                                 "not HBM; see profiles/ and DESIGN.md section 3",
                         "limiter": {"pipe": "lsu (shared-memory lookups of the fused AES-128-CTR)",
                                     "lookups_per_secret": AES_LDS_PER_SECRET,
                                     "achieved": AES_LDS_PER_SECRET * N / (share_ms * 1e-3), "peak": lds_peak,
                                     "unit": "LDS.32 lane-ops/s", "frac": AES_LDS_PER_SECRET * N / (share_ms * 1e-3) / lds_peak,
                                     "peak_source": "sclgpu_pipe_microbench(kind=4) on this GPU"}},
            "int_roofline": {"unit": "IMAD/s", "algorithmic_imads_per_secret": ALGO_IMADS_PER_SECRET,
                             "achieved": ALGO_IMADS_PER_SECRET * N / (ms_per_step * 1e-3),
                             "peak_imad32": imad_peak, "peak_imad_wide": imadw_peak,
                             "frac": ALGO_IMADS_PER_SECRET * N / (ms_per_step * 1e-3) / imad_peak,
                             "note": "peak = measured by sclgpu_pipe_microbench on this GPU just before the timed region"},
        }
        if staged is not None:
            staged["int_roofline_frac"] = ALGO_IMADS_PER_SECRET * N / (staged["ms_per_step"] * 1e-3) / imad_peak
            staged["hbm_frac_share"] = staged["share_from_coeffs_GBps"] / hbm_peak
            line["staged"] = staged
        if e2e is not None:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu_baseline:
            v, kind, cores, n_sample = cpu_reference_run(args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": v, "unit": UNIT, "cores": cores, "kind": "reference" if kind == "reference" else "port",
                "sample": f"{n_sample} secrets, SCL's per-secret shamirSecretShare+shamirRecoverP on {cores} threads "
                          f"(about {args.cpu_seconds:.0f} s of CPU work)"}
        emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
