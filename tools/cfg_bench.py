#!/usr/bin/env python
"""Device-resident timing of every BASELINE.json config (C1..C5) with CUDA events:
a development / reporting tool (bench.py is the contract; these are not bench lines).
Prints one JSON object; GB/s figures are ALGORITHMIC bytes / event time.
Usage: cfg_bench.py [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import __graft_entry__ as entry

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pkg = entry.load_package(); B = pkg.binding
ctx = pkg.Context(0); ctx.use_torch_stream()
dev = "cuda"

def timeit(fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def i64(*shape): return torch.empty(shape, dtype=torch.int64, device=dev)
res = {}

def shamir(tag, field, n, t, N, detect):
    w = 1 if field == 61 else 2
    eb = 8 * w
    sec, sh, out = i64(N, w), i64(n, N, w), i64(N, w)
    err = torch.empty(N, dtype=torch.uint8, device=dev)
    ctx.random_dev(field, "secrets", 0, N, sec)
    r = {"field": field, "n": n, "t": t, "N": N}
    r["share_ms"] = timeit(lambda: ctx.shamir_share_dev(field, sec, N, t, n, "shamir bench", 0, sh, B.PARTY_MAJOR))
    if detect:
        r["recover_d_ms"] = timeit(lambda: ctx.recover_d_dev(field, sh, N, n, t, out, err, B.PARTY_MAJOR))
        rec = r["recover_d_ms"]
        r["flags"] = int(err.sum().item())
    else:
        r["recover_p_ms"] = timeit(lambda: ctx.recover_p_dev(field, sh, N, n, out, B.PARTY_MAJOR))
        rec = r["recover_p_ms"]
    r["ok"] = bool(torch.equal(out, sec))
    if not detect and N * (t + 1) * eb <= (12 << 30):
        # staged form: coefficient planes resident in HBM (plane 0 = secrets), Polynomial::evaluate only
        planes = i64(t + 1, N, w)
        ctx.random_dev(field, "coefficient planes", 0, (t + 1) * N, planes)
        r["share_from_coeffs_ms"] = timeit(lambda: ctx.shamir_share_coeffs_dev(field, planes, N, t, n, sh, B.PARTY_MAJOR))
        r["share_from_coeffs_GBps"] = ((t + 1) * eb + n * eb) * N / r["share_from_coeffs_ms"] / 1e6
        del planes
    r["secrets_per_s"] = N / ((r["share_ms"] + rec) * 1e-3)
    r["share_GBps"] = (eb + n * eb) * N / r["share_ms"] / 1e6
    r["recover_GBps"] = ((min(n, 2 * t) if detect else n) * eb + eb) * N / rec / 1e6
    res[tag] = r
    del sec, sh, out, err

shamir("C1_fp61_n5_t2_2^20", 61, 5, 2, 1 << 20, False)
shamir("C2_fp61_n32_t15_2^26", 61, 32, 15, 1 << 26, False)
shamir("C3_fp127_n16_t7_2^24_recoverD", 127, 16, 7, 1 << 24, True)
shamir("C3b_fp61_n16_t7_2^24_recoverD", 61, 16, 7, 1 << 24, True)
torch.cuda.empty_cache()

# shamirRecoverC (Berlekamp-Welch): Fp61 n=16 (t=5), one corrupted share in every second sharing
Nc, nc_, tc_ = 1 << 20, 16, 5
sec = i64(Nc); shc = i64(nc_, Nc)
ctx.random_dev(61, "secrets", 0, Nc, sec)
ctx.shamir_share_dev(61, sec, Nc, tc_, nc_, "rc", 0, shc, B.PARTY_MAJOR)
shc[3, ::2] ^= 5
fo, eo = i64(Nc, 3 * tc_ + 1), i64(Nc, tc_ + 1)
sto = torch.empty(Nc, dtype=torch.uint8, device=dev)
r = {"field": 61, "n": nc_, "t": tc_, "N": Nc}
r["recover_c_ms"] = timeit(lambda: ctx.recover_c_dev(61, shc, Nc, nc_, fo, eo, sto, B.PARTY_MAJOR))
r["ok"] = bool(torch.equal(fo[:, 0], sec)) and int(sto.sum().item()) == 0
r["sharings_per_s"] = Nc / (r["recover_c_ms"] * 1e-3)
shc[3, ::2] ^= 5   # undo: every sharing error-free (the k_recover_c_clean path alone)
r["recover_c_error_free_ms"] = timeit(lambda: ctx.recover_c_dev(61, shc, Nc, nc_, fo, eo, sto, B.PARTY_MAJOR))
r["ok_error_free"] = bool(torch.equal(fo[:, 0], sec)) and int(sto.sum().item()) == 0
shc[3, ::64] ^= 5  # one sharing in 64 with a corrupted share
r["recover_c_1_in_64_ms"] = timeit(lambda: ctx.recover_c_dev(61, shc, Nc, nc_, fo, eo, sto, B.PARTY_MAJOR))
r["ok_1_in_64"] = bool(torch.equal(fo[:, 0], sec)) and int(sto.sum().item()) == 0
res["recoverC_fp61_n16_t5_2^20"] = r
del sec, shc, fo, eo, sto; torch.cuda.empty_cache()

# C4: PRG -> 2^28 Fp61 elements (2 GiB keystream)
n4 = 1 << 28
buf = i64(n4)
r = {"elements": n4}
r["keystream_ms"] = timeit(lambda: ctx.prg_expand_dev("prg bench", 0, 8 * n4, buf))
r["fp61_random_ms"] = timeit(lambda: ctx.random_dev(61, "prg bench", 0, n4, buf))
r["keystream_GBps"] = 8 * n4 / r["keystream_ms"] / 1e6
r["blocks_per_s"] = (n4 / 2) / (r["keystream_ms"] * 1e-3)
res["C4_prg_2^28_fp61"] = r
del buf; torch.cuda.empty_cache()

# C5: Fp61 mat-vec 8192 x 8192 and Beaver mul-add on 2^26 elements
rows = cols = 8192
A, x, y = i64(rows, cols), i64(cols), i64(rows)
ctx.random_dev(61, "mat A", 0, rows * cols, A); ctx.random_dev(61, "vec x", 0, cols, x)
r = {"rows": rows, "cols": cols}
r["matvec_ms"] = timeit(lambda: ctx.matvec_dev(61, A, rows, cols, x, y))
r["matvec_GBps"] = 8 * rows * cols / r["matvec_ms"] / 1e6
del A; torch.cuda.empty_cache()
n5 = 1 << 26
vs = [i64(n5) for _ in range(6)]
for k, v in enumerate(vs[:5]): ctx.random_dev(61, "ebdac"[k], 0, n5, v)
r["muladd_ms"] = timeit(lambda: ctx.beaver_dev(61, vs[0], vs[1], vs[2], vs[3], vs[4], n5, vs[5]))
r["muladd_GBps"] = 48 * n5 / r["muladd_ms"] / 1e6
r["vec_mul_ms"] = timeit(lambda: ctx.vec_op_dev(61, 2, vs[0], vs[1], n5, vs[5]))
r["vec_mul_GBps"] = 24 * n5 / r["vec_mul_ms"] / 1e6
r["dot_ms"] = timeit(lambda: ctx.vec_op_dev(61, 4, vs[0], vs[1], n5, vs[5]))
r["dot_GBps"] = 16 * n5 / r["dot_ms"] / 1e6
res["C5_fp61_matvec_muladd"] = r
# Matrix::multiply(Matrix), Fp61, square
del vs; torch.cuda.empty_cache()
for dim in (2048, 4096, 8192):
    A, Bm, Cm = i64(dim, dim), i64(dim, dim), i64(dim, dim)
    ctx.random_dev(61, "mat A", 0, dim * dim, A); ctx.random_dev(61, "mat B", 0, dim * dim, Bm)
    ms = timeit(lambda: ctx.matmul_dev(61, A, dim, dim, Bm, dim, Cm))
    res[f"matmul_fp61_{dim}"] = {"ms": ms, "field_mults_per_s": dim ** 3 / (ms * 1e-3), "int8_TOPS": 128 * dim ** 3 / (ms * 1e-3) / 1e12}
    del A, Bm, Cm; torch.cuda.empty_cache()
# Pedersen's sharing step: shamirSecretShare on math::Array<Fp61, 2> ({secret, randomness}), 2^25 pairs, n=32 t=15
N, W, t, n = 1 << 25, 2, 15, 32
sec, sh, out = i64(N, W), i64(n, N, W), i64(N, W)
ctx.random_dev(61, "pairs", 0, N * W, sec)
r = {"N": N, "W": W, "t": t, "n": n}
r["share_ms"] = timeit(lambda: ctx.shamir_share_array_dev(61, sec, N, W, t, n, "pedersen", 0, sh, B.PARTY_MAJOR))
r["recover_p_ms"] = timeit(lambda: ctx.recover_p_array_dev(61, sh, N, W, n, out, B.PARTY_MAJOR))
r["ok"] = bool(torch.equal(out, sec))
sm = i64(N, n, W)
r["share_secret_major_ms"] = timeit(lambda: ctx.shamir_share_array_dev(61, sec, N, W, t, n, "pedersen", 0, sm, B.SECRET_MAJOR))
r["recover_p_secret_major_ms"] = timeit(lambda: ctx.recover_p_array_dev(61, sm, N, W, n, out, B.SECRET_MAJOR))
r["ok_secret_major"] = bool(torch.equal(out, sec))
r["pairs_per_s"] = N / ((r["share_ms"] + r["recover_p_ms"]) * 1e-3)
res["array2_fp61_n32_t15_2^25_pairs"] = r
del sec, sh, out, sm

print(json.dumps(res, indent=1))
