// kernels.cuh -- the integer-pipe / HBM-bound sm_100a kernels behind libsclgpu.so (the tensor-core
// kernels are in share_tc.cu and matmul_tc.cu).
//
// One kernel per hot loop of the reference ("Kernels the new
// build must write"); each cites the reference loop it replaces.  All kernels are
// templated on the field trait (F61 / F127, field.cuh) unless a hand-specialised
// Fp61 form exists.  Unit of parallelism everywhere: one thread = one secret (or
// one element / one AES block), consecutive threads = consecutive secrets, so
// that party-major ([n][N]) share planes are read and written fully coalesced.
#pragma once
#include <cstdint>

#include "aes_ctr.cuh"
#include "aes_bitsliced.cuh"
#include "share_tc.h"
#include "field.cuh"

namespace sclgpu {

static constexpr int kAesThreads = 512;                      // 16 warps, one CTA per SM
static constexpr uint32_t kAesDynSmem = kAesTableAlign + kAesTableBytes;  // 192 KiB

// -------------------------------------------------------------------------
// common prologue of every kernel that draws keystream: build the tables, return
// this lane's table base.  The empty volatile asm pins the loads below the barrier.
__device__ __forceinline__ uint32_t aes_prologue(const uint32_t* __restrict__ g_t0) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t tbase = aes_table_base(dyn_smem);
  aes_fill_tables(tbase, g_t0);
  __syncthreads();
  uint32_t lanebase = tbase + (threadIdx.x & 31u) * 4u;
  asm volatile("" : "+r"(lanebase)::"memory");
  return lanebase;
}

// -------------------------------------------------------------------------
// Keystream blocks [first_block, first_block + n_blocks) with counter-mode caching: a
// warp takes one 256-counter group at a time (all counters sharing ctr >> 8), computes
// the group's rounds-1/2 state once (prg_group, 27 lookups) and then eight blocks per
// lane (ctr = group*256 + 32k + lane) at 133 lookups each instead of 160.  For every k
// the warp's 32 blocks are consecutive, so consumers store coalesced.
template <class Fn>
__device__ __forceinline__ void prg_for_each_block(const AesKey& key, uint32_t lanebase, uint64_t first_block,
                                                   uint64_t n_blocks, Fn&& fn) {
  if (n_blocks == 0) return;
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t g_first = first_block >> 8, g_last = (first_block + n_blocks - 1) >> 8;
  for (uint64_t grp_idx = g_first + warp; grp_idx <= g_last; grp_idx += warps) {
    PrgGroup grp;
    prg_group(key, lanebase, grp_idx << 8, grp);
#pragma unroll 1
    for (uint32_t k = 0; k < 8; ++k) {
      const uint64_t ctr = (grp_idx << 8) + 32u * k + lane;
      if (ctr >= first_block && ctr - first_block < n_blocks) {
        uint32_t o0, o1, o2, o3;
        prg_block_grouped(key, lanebase, grp, (uint32_t)ctr, o0, o1, o2, o3);
        fn(ctr - first_block, o0, o1, o2, o3);
      }
    }
  }
}

// =========================================================== PRG::next bytes
// prg.cc:124-146.  Thread = one 16-byte block at a time, written as one 128-bit store.
__global__ void __launch_bounds__(kAesThreads, 1)
k_prg_bytes(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0,
            uint64_t first_block, uint64_t n_bytes, uint8_t* __restrict__ out) {
  const uint32_t lanebase = aes_prologue(g_t0);
  const uint64_t n_full = n_bytes >> 4;
  const uint64_t n_blocks = (n_bytes + 15) >> 4;
  const bool aligned = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  prg_for_each_block(key, lanebase, first_block, n_blocks,
                     [&](uint64_t b, uint32_t o0, uint32_t o1, uint32_t o2, uint32_t o3) {
    if (b < n_full && aligned) {
      reinterpret_cast<uint4*>(out)[b] = make_uint4(o0, o1, o2, o3);
    } else {
      const uint32_t w[4] = {o0, o1, o2, o3};
      const uint64_t left = n_bytes - b * 16;
      const int m = left < 16 ? (int)left : 16;
      for (int i = 0; i < m; ++i) out[b * 16 + i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
    }
  });
}

// ===================================== Vector::random / FF::random elements
// vector.h:508-519 (PER_BLOCK = 16/BYTES elements per block, contiguous keystream)
// ff.h:72-76       (ONE_PER_BLOCK: one block per element, bytes >= BYTES dropped)
template <class F, bool ONE_PER_BLOCK>
__global__ void __launch_bounds__(kAesThreads, 1)
k_random(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0, uint64_t first_block,
         uint64_t n, typename F::E* __restrict__ out) {
  const uint32_t lanebase = aes_prologue(g_t0);
  const uint64_t per_block = (F::BYTES == 16 || ONE_PER_BLOCK) ? 1 : 2;
  const uint64_t n_blocks = (n + per_block - 1) / per_block;
  prg_for_each_block(key, lanebase, first_block, n_blocks,
                     [&](uint64_t b, uint32_t o0, uint32_t o1, uint32_t o2, uint32_t o3) {
    const uint64_t w0 = (uint64_t)o0 | ((uint64_t)o1 << 32), w1 = (uint64_t)o2 | ((uint64_t)o3 << 32);
    if constexpr (F::BYTES == 16) {
      out[b] = F127::from_raw(E127{w0, w1});
    } else if constexpr (ONE_PER_BLOCK) {
      out[b] = F61::from_raw(w0);
    } else {
      const uint64_t e0 = F61::from_raw(w0), e1 = F61::from_raw(w1);
      if (2 * b + 1 < n && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        reinterpret_cast<ulonglong2*>(out)[b] = make_ulonglong2(e0, e1);
      } else {
        out[2 * b] = e0;
        if (2 * b + 1 < n) out[2 * b + 1] = e1;
      }
    }
  });
}

// FF::read on packed bytes (ff.h:63-67): n elements, bytes little-endian
template <class F>
__global__ void k_from_bytes(const uint8_t* __restrict__ bytes, uint64_t n,
                             typename F::E* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint8_t* p = bytes + i * F::BYTES;
    uint64_t w[2] = {0, 0};
#pragma unroll
    for (int k = 0; k < F::BYTES; ++k) w[k >> 3] |= (uint64_t)p[k] << (8 * (k & 7));
    if constexpr (F::BYTES == 16) {
      out[i] = F127::from_raw(E127{w[0], w[1]});
    } else {
      out[i] = F61::from_raw(w[0]);
    }
  }
}

// ======================================================== Horner, small point
// Polynomial::evaluate (poly.h:56-64): y = c_k + y*x from the top coefficient.
// The evaluation points of shamirSecretShare are 1..n (shamir.h:62-65), i.e. small
// integers, so one Horner step is a (field element) x (small integer) product.
//
// Fp61: y is kept semi-reduced (y < 2^61 + 2^18) in two 32-bit limbs; for x < 2^16
//   u = y0*x + c            (IMAD.WIDE, 64-bit addend)            < 2^63
//   v = y1*x + (u >> 32)    (IMAD.WIDE)                           < 2^47
//   y' = (u mod 2^32) + ((v mod 2^29) << 32) + (v >> 29)          2^61 = 1 (mod p)
template <class F>
struct Horner;

template <>
struct Horner<F61> {
  typedef uint64_t Y;
  static constexpr uint32_t MAX_SMALL_X = 0xFFFF;
  static SCLGPU_D Y init(uint64_t c) { return c; }
  static SCLGPU_D Y step(Y y, uint32_t x, uint64_t c) {
    const uint32_t y0 = (uint32_t)y, y1 = (uint32_t)(y >> 32);
    const uint64_t u = (uint64_t)y0 * x + c;
    const uint64_t v = (uint64_t)y1 * x + (u >> 32);
    return ((u & 0xFFFFFFFFULL) | ((v & 0x1FFFFFFFULL) << 32)) + (v >> 29);
  }
  static SCLGPU_D uint64_t finish(Y y) {  // y < 2^62
    uint64_t r = (y & F61::P) + (y >> 61);
    return r >= F61::P ? r - F61::P : r;
  }
};

// Fp127: canonical after every step.  y*x + c with x < 2^16:
//   (l0,l1) = y.lo*x ; (h0,h1) = y.hi*x + l1 ; value = l0 + h0*2^64 + h1*2^128, 2^127 = 1
template <>
struct Horner<F127> {
  typedef E127 Y;
  static constexpr uint32_t MAX_SMALL_X = 0xFFFF;
  static SCLGPU_D Y init(E127 c) { return c; }
  static SCLGPU_D Y step(Y y, uint32_t x, E127 c) {
    uint64_t l0, l1, h0, h1;
    mul64wide(y.lo, x, l0, l1);
    mul64wide(y.hi, x, h0, h1);
    h0 += l1;
    h1 += (h0 < l1);
    const uint64_t top = (h0 >> 63) | (h1 << 1);  // < 2^18
    E127 r{l0, h0 & F127::PHI};                    // < 2^127
    r.lo += top;
    r.hi += (r.lo < top);                          // <= 2^127 + small: still < 2^128
    r = F127::from_raw(r);
    return F127::add(r, c);
  }
  static SCLGPU_D E127 finish(Y y) { return y; }
};

// generic-point fallback (x is a field element, any size): plain mul + add
template <class F>
__device__ __forceinline__ typename F::E horner_generic(const typename F::E* c, int t,
                                                        typename F::E x) {
  typename F::E y = c[t];
  for (int k = t - 1; k >= 0; --k) y = F::add(c[k], F::mul(y, x));
  return y;
}

// ============================================== shamirSecretShare, fused with PRG
// shamir.h:52-68 + vector.h:508-519 + prg.cc:124-146, N calls on one PRG.
// Thread = one secret.  Its B = ceil((T+1)*BYTES/16) keystream blocks start at
// first_block + j*B; coefficient k (1..T) is word k of that keystream (slot 0 is
// drawn and replaced by the secret).  Coefficients live in registers; the n
// evaluations are independent Horner chains (ILP), each result stored coalesced
// into share plane i (out[i*stride_i + j*stride_j]).
template <class F, int T>
__global__ void __launch_bounds__(kAesThreads, 1)
k_share_fused(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0,
              uint64_t first_block, const typename F::E* __restrict__ secrets, uint64_t N,
              uint32_t n, typename F::E* __restrict__ out, uint64_t stride_i, uint64_t stride_j) {
  typedef typename F::E E;
  typedef Horner<F> H;
  const uint32_t lanebase = aes_prologue(g_t0);
  constexpr uint64_t B = ((uint64_t)(T + 1) * F::BYTES + 15) / 16;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    E c[T + 1];
    c[0] = secrets[j];
    const uint64_t ctr0 = first_block + j * B;
    if constexpr (F::BYTES == 8) {
#pragma unroll
      for (int b = 0; b < (int)B; ++b) {
        uint32_t o0, o1, o2, o3;
        prg_block(key, lanebase, ctr0 + b, o0, o1, o2, o3);
        if (2 * b >= 1 && 2 * b <= T) c[2 * b] = F61::from_raw((uint64_t)o0 | ((uint64_t)o1 << 32));
        if (2 * b + 1 <= T) c[2 * b + 1] = F61::from_raw((uint64_t)o2 | ((uint64_t)o3 << 32));
      }
    } else {
#pragma unroll
      for (int b = 1; b <= T; ++b) {  // block 0 = slot 0: consumed, never used
        uint32_t o0, o1, o2, o3;
        prg_block(key, lanebase, ctr0 + b, o0, o1, o2, o3);
        c[b] = F127::from_raw(E127{(uint64_t)o0 | ((uint64_t)o1 << 32), (uint64_t)o2 | ((uint64_t)o3 << 32)});
      }
    }
    E* dst = out + j * stride_j;
    if (n <= H::MAX_SMALL_X) {
#pragma unroll 4
      for (uint32_t i = 1; i <= n; ++i) {
        typename H::Y y = H::init(c[T]);
#pragma unroll
        for (int k = T - 1; k >= 0; --k) y = H::step(y, i, c[k]);
        dst[(uint64_t)(i - 1) * stride_i] = H::finish(y);
      }
    } else {
      for (uint32_t i = 1; i <= n; ++i) {
        dst[(uint64_t)(i - 1) * stride_i] = horner_generic<F>(c, T, F::from_u32(i));
      }
    }
  }
}

// ================================= shamirSecretShare, Fp61, tuned (the C2 kernel)
// Same contract as k_share_fused<F61, T>; differences (DESIGN.md "share kernel"):
//  * AES blocks are produced one at a time in a rolled loop with counter-mode
//    caching (prg_group / prg_block_grouped) and staged through shared memory
//    ([k][thread] u64, conflict-free), so the code stays small (no instruction
//    cache misses) and the coefficients are re-read into registers only once;
//  * the Horner step is 5 instructions (h61_step): 2 IMAD.WIDE, one shift and an
//    IADD3 / IADD3.X pair;
//  * four evaluation points are interleaved per rolled iteration (ILP 4).
// State invariant: y = y0 + y1*2^32 < 2^63 is congruent to the Horner value.
// One Horner step y <- y*x + c (mod p), x < 2^16, x8 = 8*x.
//   u  = y0*x + c                     (IMAD.WIDE)          < 2^61 + 2^48
//   v8 = y1*(8x) = 8*(y1*x)           (IMAD.WIDE)          < 2^50
//   y1*x*2^32 = vl*2^32 + vh*2^61 with vh = v8 >> 32, vl = (v8 mod 2^32) >> 3, and 2^61 = 1
//   y' = u + vh + vl*2^32             (IADD3, IADD3.X)     < 2^62 + 2^49
// VLMODE selects how vl is formed: 0 = shift (ALU pipe), 1 = mul.hi by 2^29 (FMA
// pipe), 2 = alternate by coefficient index -- a pipe-balancing knob only.
template <int VLMODE>
__device__ __forceinline__ void h61_step(uint32_t& y0, uint32_t& y1, uint32_t x, uint32_t x8, uint64_t c,
                                         uint32_t two29, int k) {
  uint64_t pr, v8;
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(pr) : "r"(y0), "r"(x));
  asm("mul.wide.u32 %0, %1, %2;" : "=l"(v8) : "r"(y1), "r"(x8));
  const uint32_t vlo = (uint32_t)v8, vh = (uint32_t)(v8 >> 32);
  uint32_t vl;
  if (VLMODE == 1 || (VLMODE == 2 && (k & 1))) {
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(vl) : "r"(vlo), "r"(two29));
  } else {
    vl = vlo >> 3;
  }
  const uint64_t u = pr + c;  // ptxas folds this into the IMAD.WIDE addend
  asm("add.cc.u32 %0, %2, %3;\n\taddc.u32 %1, %4, %5;"
      : "=r"(y0), "=r"(y1)
      : "r"((uint32_t)u), "r"(vh), "r"((uint32_t)(u >> 32)), "r"(vl));
}

__device__ __forceinline__ uint64_t h61_finish(uint32_t y0, uint32_t y1) {
  const uint64_t y = (uint64_t)y0 | ((uint64_t)y1 << 32);
  const uint64_t r = (y & F61::P) + (y >> 61);
  return r >= F61::P ? r - F61::P : r;
}

static constexpr uint32_t kShare61StageStride = kAesThreads * 8;  // bytes per coefficient row

template <int T, int VLMODE>
__global__ void __launch_bounds__(kAesThreads, 1)
k_share61(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0, uint64_t first_block,
          const uint64_t* __restrict__ secrets, uint64_t N, uint32_t n, uint64_t* __restrict__ out,
          uint64_t stride_i, uint64_t stride_j, uint32_t two29) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t dyn = smem_u32(dyn_smem);
  const uint32_t tbase = aes_table_base(dyn_smem);
  constexpr uint32_t kStageBytes = (T > 0 ? T : 1) * kShare61StageStride;
  // the 64 KiB alignment slack of the tables is either before or after them
  uint32_t stage = (dyn + 15u) & ~15u;
  if (tbase - stage < kStageBytes) stage = tbase + kAesTableBytes;
  if (stage + kStageBytes > dyn + kAesDynSmem) __trap();  // cannot happen with <= 4 KiB of static smem
  aes_fill_tables(tbase, g_t0);
  __syncthreads();
  uint32_t lanebase = tbase + (threadIdx.x & 31u) * 4u;
  asm volatile("" : "+r"(lanebase)::"memory");
  const uint32_t my_stage = stage + threadIdx.x * 8u;

  constexpr uint32_t B = ((uint32_t)(T + 1) * 8 + 15) / 16;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    const uint64_t ctr0 = first_block + j * B;
    if (T > 0) {
      PrgGroup g;
      uint64_t gid = ctr0 >> 8;
      prg_group(key, lanebase, ctr0, g);
#pragma unroll 1
      for (uint32_t b = 0; b < B; ++b) {
        const uint64_t ctr = ctr0 + b;
        if ((ctr >> 8) != gid) {  // crossed a 256-block group: rare, at most once per secret
          gid = ctr >> 8;
          prg_group(key, lanebase, ctr, g);
        }
        uint32_t o0, o1, o2, o3;
        prg_block_grouped(key, lanebase, g, (uint32_t)ctr, o0, o1, o2, o3);
        // coefficient 2b (b >= 1) and 2b+1 (<= T); row k-1 of the stage
        const uint32_t a = my_stage + (2 * b) * kShare61StageStride;
        if (b >= 1) asm volatile("st.shared.v2.u32 [%0+%3], {%1, %2};" ::"r"(a), "r"(o0), "r"(o1), "n"(-(int)kShare61StageStride) : "memory");
        if (2 * b + 1 <= (uint32_t)T) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(o2), "r"(o3) : "memory");
      }
    }
    uint64_t c[T + 1];
    c[0] = secrets[j];
    asm volatile("" : "+l"(c[0]));
#pragma unroll
    for (int k = 1; k <= T; ++k) {
      uint32_t lo, hi;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(my_stage + (uint32_t)(k - 1) * kShare61StageStride) : "memory");
      // semi-reduced (<= p + 7) is enough: the Horner state is lazy and h61_finish canonicalises
      c[k] = ((uint64_t)lo | ((uint64_t)(hi & 0x1FFFFFFFu) << 32)) + (hi >> 29);
      asm volatile("" : "+l"(c[k]));  // keep it one 64-bit register pair: the IMAD.WIDE addend
    }
    uint64_t* d = out + j * stride_j;
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < n; i0 += 4, d += 4 * stride_i) {
      uint32_t ya0 = (uint32_t)c[T], ya1 = (uint32_t)(c[T] >> 32);
      uint32_t yb0 = ya0, yb1 = ya1, yc0 = ya0, yc1 = ya1, yd0 = ya0, yd1 = ya1;
      uint32_t xa = i0 + 1;
      asm volatile("" : "+r"(xa));  // keep the point a plain 32-bit register value
      const uint32_t xb = xa + 1, xc = xa + 2, xd = xa + 3;
      const uint32_t xa8 = xa << 3, xb8 = xb << 3, xc8 = xc << 3, xd8 = xd << 3;
#pragma unroll
      for (int k = T - 1; k >= 0; --k) {
        h61_step<VLMODE>(ya0, ya1, xa, xa8, c[k], two29, k);
        h61_step<VLMODE>(yb0, yb1, xb, xb8, c[k], two29, k);
        h61_step<VLMODE>(yc0, yc1, xc, xc8, c[k], two29, k);
        h61_step<VLMODE>(yd0, yd1, xd, xd8, c[k], two29, k);
      }
      d[0] = h61_finish(ya0, ya1);
      if (i0 + 1 < n) d[stride_i] = h61_finish(yb0, yb1);
      if (i0 + 2 < n) d[2 * stride_i] = h61_finish(yc0, yc1);
      if (i0 + 3 < n) d[3 * stride_i] = h61_finish(yd0, yd1);
    }
  }
}

static constexpr uint8_t kRecoverCPending = 2;
static constexpr uint32_t kRecoverCMaxT = 10;  // 3t+1 <= 32

// Error-free sharings, one thread each.  If all np = 3t+1 points lie on one polynomial f of degree <= t, every
// Berlekamp-Welch system with e >= 1 has many solutions (E any monic polynomial of degree e, Q = f*E), so the
// reference's solveLinearSystem rejects e = t..1 (matrix.h:741-764, unique solutions only) and accepts e = 0, whose
// unique solution is Q = f, E = 1: the result is the interpolant's coefficients and the error locator (1).
// check: 2t Lagrange rows (nodes a_0..a_t evaluated at a_{t+1}..a_{3t}), coef: the (t+1) x (t+1) matrix taking the
// first t+1 shares to f's coefficients.  Sharings that fail a check are listed in `pending` for k_recover_c.
template <class F>
__global__ void __launch_bounds__(256)
k_recover_c_clean(const typename F::E* __restrict__ in, uint64_t N, uint64_t stride_i, uint64_t stride_j,
                  uint32_t t, const typename F::E* __restrict__ check, const typename F::E* __restrict__ coef,
                  typename F::E* __restrict__ f_out, typename F::E* __restrict__ e_out,
                  uint8_t* __restrict__ status, uint32_t* __restrict__ pending,
                  unsigned long long* __restrict__ n_pending) {
  typedef typename F::E E;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  E* sm = reinterpret_cast<E*>(dyn_smem);
  const uint32_t m = t + 1u, np = 3u * t + 1u, n_rows = 3u * t + 1u;  // 2t check rows, then t+1 coefficient rows
  for (uint32_t i = threadIdx.x; i < n_rows * m; i += blockDim.x) sm[i] = i < 2u * t * m ? check[i] : coef[i - 2u * t * m];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    const E* src = in + j * stride_j;
    E s[kRecoverCMaxT + 1];
#pragma unroll
    for (uint32_t k = 0; k <= kRecoverCMaxT; ++k) s[k] = k < m ? src[(uint64_t)k * stride_i] : F::zero();
    bool clean = true;
    E* fo = f_out + j * np;
    for (uint32_t r = 0; r < n_rows; ++r) {
      const E* row = sm + r * m;
      typename F::Acc acc = F::acc_zero();
#pragma unroll
      for (uint32_t k = 0; k <= kRecoverCMaxT; ++k)
        if (k < m) F::mac(acc, s[k], row[k]);  // m <= 11 terms: no fold needed
      const E y = F::acc_reduce(acc);
      if (r < 2u * t) {
        clean = clean && F::eq(y, src[(uint64_t)(m + r) * stride_i]);
      } else if (clean) {
        fo[r - 2u * t] = y;
      }
      if (r + 1u == 2u * t && !clean) break;
    }
    if (clean) {
      for (uint32_t k = m; k < np; ++k) fo[k] = F::zero();
      E* eo = e_out + j * (uint64_t)m;
      eo[0] = F::one();
      for (uint32_t k = 1; k < m; ++k) eo[k] = F::zero();
      status[j] = 0;
    } else {
      status[j] = kRecoverCPending;
      pending[atomicAdd(n_pending, 1ull)] = (uint32_t)j;  // compacted: the elimination kernel stays load-balanced
    }
  }
}

// ================================================================ shamirRecoverC
// include/scl/ss/shamir.h:203-258 (Berlekamp-Welch) + solveLinearSystem (matrix.h:812-828) +
// Polynomial::divide (poly.h:262-278).  One WARP per sharing, lane i owns row i of the
// np x (np+1) augmented system in shared memory (np = 3t+1 <= 32).  For e = t..0:
//   A(i,j) = s_i a_i^j (j<e), A(i,e) = -1, A(i,j) = A(i,j-1) a_i (j>e), b_i = -s_i a_i^e
// The reference accepts a system only when its solution is unique (hasSolution(aug, true),
// matrix.h:741-764), so any exact elimination yields the same x.  Here: fraction-free
// Gauss-Jordan (row_k <- p*row_k - f*row_c: no inversion per pivot), then one inversion per
// lane, all lanes in parallel.  Then Q / E by long division, lanes over the divisor terms.
// Outputs per sharing: f (np coefficients, zero padded), err (t+1, monic, zero padded),
// status 1 where the reference throws "could not correct shares" (f, err zeroed).
// quick != 0: first try the verified one-elimination shortcut described in the loop body.
template <class F>
__global__ void __launch_bounds__(256)
k_recover_c(const typename F::E* __restrict__ in, uint64_t N, uint64_t stride_i, uint64_t stride_j,
            uint32_t t, const typename F::E* __restrict__ alphas, typename F::E* __restrict__ f_out,
            typename F::E* __restrict__ e_out, uint8_t* __restrict__ status,
            unsigned long long* __restrict__ n_failed, const uint32_t* __restrict__ pending,
            const unsigned long long* __restrict__ n_pending, int quick) {
  typedef typename F::E E;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t np = 3u * t + 1u, cols = np + 1u;
  const uint32_t per_warp = np * cols + 3u * np;            // matrix, x, division remainder, quotient
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  E* M = reinterpret_cast<E*>(dyn_smem) + (size_t)wib * per_warp;
  E* X = M + np * cols;
  E* R = X + np;
  E* Qt = R + np;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const E minus1 = F::neg(F::one());
  const bool row = lane < np;
  const E a_i = row ? alphas[lane] : F::zero();
  unsigned long long local_failed = 0;

  const uint64_t n_work = pending ? *n_pending : N;   // with a list: only what k_recover_c_clean left over
  for (uint64_t q = warp; q < n_work; q += warps) {
    const uint64_t j = pending ? pending[q] : q;
    const E s_i = row ? in[(uint64_t)lane * stride_i + j * stride_j] : F::zero();
    if (quick && t > 0) {
      // ---- one elimination instead of up to t+1.  ANY solution (E, Q) of the e = t system gives the decoded
      // polynomial f = Q / E when the word is within t errors of a codeword (Berlekamp-Welch).  So: rank-revealing
      // fraction-free Gauss-Jordan without row exchanges, free unknowns := 0, divide, and VERIFY: if f has degree
      // <= t and disagrees with the shares in d <= t places, unique decoding makes the reference's answer the
      // unique solution of its e = d system (every e > d system has many solutions and is rejected), namely f and
      // the locator prod_{bad i} (x - a_i) -- written here directly.  Anything else falls through to the
      // reference's own sequence below.
      if (row) {
        E* mr = M + lane * cols;
        E v = s_i;
        for (uint32_t c = 0; c < t; ++c) {
          mr[c] = v;
          v = F::mul(v, a_i);
        }
        mr[np] = F::neg(v);
        E u = minus1;
        for (uint32_t c = t; c < np; ++c) {
          mr[c] = u;
          u = F::mul(u, a_i);
        }
        X[lane] = F::zero();
      }
      __syncwarp();
      // np <= 16: two lanes per row (lane and lane + 16), each updating every other column
      const bool two = np <= 16u;
      const uint32_t sub = two ? (lane >> 4) : 0u, rl = two ? (lane & 15u) : lane, qs = two ? 2u : 1u;
      const bool row2 = rl < np;
      int my_col = -1;  // pivot column of this lane's row
      for (uint32_t c = 0; c < np; ++c) {
        const bool nz = row && my_col < 0 && !F::is_zero(M[lane * cols + c]);
        const unsigned mask = __ballot_sync(0xffffffffu, nz);
        if (mask == 0) continue;  // free unknown
        const uint32_t piv = (uint32_t)__ffs(mask) - 1u;
        if (row2 && rl == piv) my_col = (int)c;
        const E p = M[piv * cols + c];
        const bool upd = row2 && rl != piv;
        E* mr = M + rl * cols;
        const E f = upd ? mr[c] : F::zero();
        __syncwarp();  // both lanes of a row have read column c before it is cleared
        if (upd && !F::is_zero(f)) {
          const E nf = F::neg(f);  // row*p - f*pivot_row as ONE lazily reduced two-term sum
          for (uint32_t q = c + 1 + sub; q < cols; q += qs) {
            typename F::Acc acc = F::acc_zero();
            F::mac(acc, mr[q], p);
            F::mac(acc, nf, M[piv * cols + q]);
            mr[q] = F::acc_reduce(acc);
          }
          if (sub == 0) {
            mr[c] = F::zero();
            if (my_col >= 0) mr[my_col] = F::mul(mr[my_col], p);
          }
        }
        __syncwarp();
      }
      // rows without a pivot read 0 = b: consistent iff b == 0
      bool good = __ballot_sync(0xffffffffu, row && my_col < 0 && !F::is_zero(M[lane * cols + np])) == 0;
      if (good) {
        if (row && my_col >= 0) X[my_col] = F::mul(M[lane * cols + np], F::inv(M[lane * cols + my_col]));
        __syncwarp();
        const uint32_t qn = np - t;  // Q = x[t..np-1], E = (x_0..x_{t-1}, 1)
        if (lane < qn) R[lane] = X[t + lane];
        if (lane < np) Qt[lane] = F::zero();
        __syncwarp();
        const unsigned nzq = __ballot_sync(0xffffffffu, lane < qn && !F::is_zero(R[lane]));
        const uint32_t deg = nzq ? 31u - (uint32_t)__clz(nzq) : 0u;
        if (deg >= t) {
          for (int d = (int)deg; d >= (int)t; --d) {
            const E c = R[d];
            __syncwarp();
            if (lane < t) R[d - t + lane] = F::sub(R[d - t + lane], F::mul(c, X[lane]));
            if (lane == 0) {
              Qt[d - t] = c;
              R[d] = F::zero();
            }
            __syncwarp();
          }
          good = __ballot_sync(0xffffffffu, lane < t && !F::is_zero(R[lane])) == 0;
        } else {
          good = nzq == 0;
        }
        // f = Qt has degree <= deg - t <= t by construction; count the disagreements with the shares
        E y = F::zero();
        if (row) {
          for (int k = (int)t; k >= 0; --k) y = F::add(F::mul(y, a_i), Qt[k]);
        }
        const unsigned badm = __ballot_sync(0xffffffffu, row && !F::eq(y, s_i));
        const uint32_t d_err = (uint32_t)__popc(badm);
        if (good && d_err <= t) {
          E* fo = f_out + j * np;
          E* eo = e_out + j * (uint64_t)(t + 1);
          if (lane < np) fo[lane] = Qt[lane];
          // locator, built in R by lane 0: prod over the bad positions of (x - a_i), low coefficient first
          if (lane == 0) {
            R[0] = F::one();
            uint32_t dg = 0;
            for (unsigned mm = badm; mm; mm &= mm - 1u) {
              const E a = alphas[__ffs(mm) - 1];
              R[dg + 1] = R[dg];
              for (uint32_t q = dg; q >= 1; --q) R[q] = F::sub(R[q - 1], F::mul(a, R[q]));
              R[0] = F::neg(F::mul(a, R[0]));
              ++dg;
            }
            for (uint32_t q = 0; q <= t; ++q) eo[q] = q <= dg ? R[q] : F::zero();
            status[j] = 0;
          }
          __syncwarp();
          continue;
        }
      }
      __syncwarp();
    }
    int e = (int)t;
    for (;; --e) {
      if (row) {
        E* mr = M + lane * cols;
        E v = s_i;
        for (int c = 0; c < e; ++c) {
          mr[c] = v;
          v = F::mul(v, a_i);
        }
        mr[np] = F::neg(v);
        E u = minus1;
        for (uint32_t c = (uint32_t)e; c < np; ++c) {
          mr[c] = u;
          u = F::mul(u, a_i);
        }
      }
      __syncwarp();
      bool singular = false;
      for (uint32_t c = 0; c < np; ++c) {
        const bool nz = row && lane >= c && !F::is_zero(M[lane * cols + c]);
        const unsigned mask = __ballot_sync(0xffffffffu, nz);
        if (mask == 0) {
          singular = true;
          break;
        }
        const uint32_t piv = (uint32_t)__ffs(mask) - 1u;
        if (piv != c) {
          for (uint32_t q = lane; q < cols; q += 32u) {
            const E tmp = M[piv * cols + q];
            M[piv * cols + q] = M[c * cols + q];
            M[c * cols + q] = tmp;
          }
          __syncwarp();
        }
        const E p = M[c * cols + c];
        if (row && lane != c) {
          E* mr = M + lane * cols;
          const E f = mr[c];
          if (!F::is_zero(f)) {
            const E nf = F::neg(f);
            for (uint32_t q = c + 1; q < cols; ++q) {
              typename F::Acc acc = F::acc_zero();
              F::mac(acc, mr[q], p);
              F::mac(acc, nf, M[c * cols + q]);
              mr[q] = F::acc_reduce(acc);
            }
            mr[c] = F::zero();
            if (lane < c) mr[lane] = F::mul(mr[lane], p);  // keep the diagonal of finished rows consistent
          }
        }
        __syncwarp();
      }
      if (!singular) break;
      if (e == 0) {  // only with coinciding nodes: the reference then proceeds with x = 0 (shamir.h:213,238-240)
        e = -1;
        break;
      }
      __syncwarp();
    }
    // x_i = b_i / d_i
    if (row) X[lane] = e < 0 ? F::zero() : F::mul(M[lane * cols + np], F::inv(M[lane * cols + lane]));
    if (e < 0) e = 0;
    __syncwarp();
    // Q = x[e..np-1] (trailing zeros stripped), E = (x_0..x_{e-1}, 1); f = Q / E
    const uint32_t ue = (uint32_t)e, qn = np - ue;
    if (lane < qn) R[lane] = X[ue + lane];
    __syncwarp();
    const unsigned nzq = __ballot_sync(0xffffffffu, lane < qn && !F::is_zero(R[lane]));
    const uint32_t deg = nzq ? 31u - (uint32_t)__clz(nzq) : 0u;
    E* fo = f_out + j * np;
    E* eo = e_out + j * (uint64_t)(t + 1);
    if (lane < np) fo[lane] = F::zero();
    if (lane <= t) eo[lane] = F::zero();
    __syncwarp();
    bool ok;
    if (deg >= ue) {
      for (int d = (int)deg; d >= (int)ue; --d) {
        const E c = R[d];
        __syncwarp();
        if (lane < ue) R[d - ue + lane] = F::sub(R[d - ue + lane], F::mul(c, X[lane]));
        if (lane == 0) {
          fo[d - ue] = c;
          R[d] = F::zero();
        }
        __syncwarp();
      }
      ok = __ballot_sync(0xffffffffu, lane < ue && !F::is_zero(R[lane])) == 0;
    } else {
      ok = nzq == 0;  // Q == 0: quotient and remainder are zero
    }
    if (ok) {
      if (lane < ue) eo[lane] = X[lane];
      if (lane == ue) eo[lane] = F::one();
      if (lane == 0) status[j] = 0;
    } else {
      __syncwarp();
      if (lane < np) fo[lane] = F::zero();
      if (lane == 0) {
        status[j] = 1;
        ++local_failed;
      }
    }
    __syncwarp();
  }
  if (local_failed) atomicAdd(n_failed, local_failed);
}

// ================================================================ additiveShare
// include/scl/ss/additive.h:42-53, N calls on one PRG: n-1 x FF::random (ff.h:72-76: ONE
// whole keystream block per element, bytes beyond byteSize dropped), last share =
// secret - sum.  Secret j draws blocks [first_block + j(n-1), first_block + (j+1)(n-1)).
// Thread = one secret; share (j,i) at out[i*stride_i + j*stride_j].  n >= 1.
template <class F>
__global__ void __launch_bounds__(kAesThreads, 1)
k_additive_share(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0, uint64_t first_block,
                 const typename F::E* __restrict__ secrets, uint64_t N, uint32_t n,
                 typename F::E* __restrict__ out, uint64_t stride_i, uint64_t stride_j) {
  typedef typename F::E E;
  const uint32_t lanebase = aes_prologue(g_t0);
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint32_t draws = n - 1u;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    E* dst = out + j * stride_j;
    E sum = F::zero();
    if (draws) {
      const uint64_t ctr0 = first_block + j * draws;
      PrgGroup grp;
      uint64_t gid = ctr0 >> 8;
      prg_group(key, lanebase, ctr0, grp);
#pragma unroll 1
      for (uint32_t i = 0; i < draws; ++i) {
        const uint64_t ctr = ctr0 + i;
        if ((ctr >> 8) != gid) {
          gid = ctr >> 8;
          prg_group(key, lanebase, ctr, grp);
        }
        uint32_t o0, o1, o2, o3;
        prg_block_grouped(key, lanebase, grp, (uint32_t)ctr, o0, o1, o2, o3);
        E s;
        if constexpr (F::BYTES == 16) {
          s = F127::from_raw(E127{(uint64_t)o0 | ((uint64_t)o1 << 32), (uint64_t)o2 | ((uint64_t)o3 << 32)});
        } else {
          s = F61::from_raw((uint64_t)o0 | ((uint64_t)o1 << 32));
        }
        dst[(uint64_t)i * stride_i] = s;
        sum = F::add(sum, s);
      }
    }
    dst[(uint64_t)draws * stride_i] = F::sub(secrets[j], sum);
  }
}

// reconstruction of additive sharings = Vector::sum per sharing (vector.h:262-267)
template <class F>
__global__ void __launch_bounds__(256)
k_additive_recover(const typename F::E* __restrict__ in, uint64_t N, uint32_t n, uint64_t stride_i,
                   uint64_t stride_j, typename F::E* __restrict__ out) {
  typedef typename F::E E;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    const E* src = in + j * stride_j;
    E sum = F::zero();
    uint32_t i = 0;
    for (; i + 8 <= n; i += 8) {
      E v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = src[(uint64_t)(i + k) * stride_i];
#pragma unroll
      for (int k = 0; k < 8; ++k) sum = F::add(sum, v[k]);
    }
    for (; i < n; ++i) sum = F::add(sum, src[(uint64_t)i * stride_i]);
    out[j] = sum;
  }
}

// PRG -> coefficient planes, for thresholds without a register-resident
// instantiation and for array-valued secrets (shamirSecretShare on
// math::Array<FF, W>, the sharing step of pedersen.h:137-138).  Sharing j draws
// one Vector<Array>::random(t+1): stream element s = k*W + w is component w of
// coefficient k (array.h:82-88 inside vector.h:508-519), B = ceil((t+1)*W*bs/16)
// blocks per sharing; elements s < W are consumed and replaced by the secret.
// Output: coeffs[k*(N*W) + j*W + w], i.e. N*W independent polynomials.  W = 1
// is the plain shamirSecretShare.
template <class F>
__global__ void __launch_bounds__(kAesThreads, 1)
k_expand_coeffs(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0,
                uint64_t first_block, const typename F::E* __restrict__ secrets, uint64_t N,
                uint32_t t, uint32_t W, typename F::E* __restrict__ coeffs) {
  const uint32_t lanebase = aes_prologue(g_t0);
  const uint64_t S = (uint64_t)(t + 1) * W;  // stream elements per sharing
  const uint64_t B = (S * F::BYTES + 15) / 16;
  const uint64_t NW = N * W;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  auto put = [&](uint64_t j, uint64_t s, typename F::E v) {
    if (s >= W && s < S) coeffs[(s / W) * NW + j * W + (s % W)] = v;
  };
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    for (uint32_t w = 0; w < W; ++w) coeffs[j * W + w] = secrets[j * W + w];
    const uint64_t ctr0 = first_block + j * B;
    PrgGroup grp;  // rounds 1-2 of the 256-block group of the counter, cached (aes_ctr.cuh)
    uint64_t gid = ~0ull;
    for (uint64_t b = (F::BYTES == 16 ? W : W / 2); b < B; ++b) {
      uint32_t o0, o1, o2, o3;
      const uint64_t ctr = ctr0 + b;
      if ((ctr >> 8) != gid) {
        gid = ctr >> 8;
        prg_group(key, lanebase, ctr, grp);
      }
      prg_block_grouped(key, lanebase, grp, (uint32_t)ctr, o0, o1, o2, o3);
      const uint64_t w0 = (uint64_t)o0 | ((uint64_t)o1 << 32), w1 = (uint64_t)o2 | ((uint64_t)o3 << 32);
      if constexpr (F::BYTES == 16) {
        put(j, b, F127::from_raw(E127{w0, w1}));
      } else {
        put(j, 2 * b, F61::from_raw(w0));
        put(j, 2 * b + 1, F61::from_raw(w1));
      }
    }
  }
}

// Polynomial evaluation from coefficient planes (poly.h:56-64): any t, any n.
template <class F>
__global__ void __launch_bounds__(256)
k_share_coeffs(const typename F::E* __restrict__ coeffs, uint64_t N, uint32_t t, uint32_t n,
               typename F::E* __restrict__ out, uint64_t stride_i, uint64_t stride_j) {
  typedef typename F::E E;
  typedef Horner<F> H;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    E* dst = out + j * stride_j;
    for (uint32_t i = 1; i <= n; ++i) {
      E r;
      if (n <= H::MAX_SMALL_X) {
        typename H::Y y = H::init(coeffs[(uint64_t)t * N + j]);
        for (int64_t k = (int64_t)t - 1; k >= 0; --k) y = H::step(y, i, coeffs[(uint64_t)k * N + j]);
        r = H::finish(y);
      } else {
        const E x = F::from_u32(i);
        r = coeffs[(uint64_t)t * N + j];
        for (int64_t k = (int64_t)t - 1; k >= 0; --k) r = F::add(coeffs[(uint64_t)k * N + j], F::mul(r, x));
      }
      dst[(uint64_t)(i - 1) * stride_i] = r;
    }
  }
}

// ================================================== computeLagrangeBasis rows
// lagrange.h:55-71: row r, entry i = prod_{j != i} (x_r - nodes[j]) / (nodes[i] - nodes[j]).
// grid = rows, block >= m threads.  xs[r] is the evaluation point of row r.
// *bad is set when a denominator is zero ("0 not invertible modulo prime").
template <class F>
__global__ void k_lagrange_rows(const typename F::E* __restrict__ nodes, uint32_t m,
                                const typename F::E* __restrict__ xs,
                                typename F::E* __restrict__ out, int* bad) {
  typedef typename F::E E;
  const E x = xs[blockIdx.x];
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) {
    const E xi = nodes[i];
    E ell = F::one();
    for (uint32_t j = 0; j < m; ++j) {
      if (j == i) continue;
      const E den = F::sub(xi, nodes[j]);
      if (F::is_zero(den)) {
        atomicExch(bad, 1);
        continue;
      }
      ell = F::mul(ell, F::mul(F::sub(x, nodes[j]), F::inv(den)));
    }
    out[(uint64_t)blockIdx.x * m + i] = ell;
  }
}

// ============================================================ shamirRecoverP
// shamir.h:82-87: out[j] = sum_i shares[j][i] * basis[i]  (innerProd, vector.h:45-52)
// share (j,i) at in[i*stride_i + j*stride_j]; basis staged in shared memory.
template <class F>
__global__ void __launch_bounds__(256)
k_recover_p(const typename F::E* __restrict__ in, uint64_t N, uint32_t n, uint64_t stride_i,
            uint64_t stride_j, const typename F::E* __restrict__ basis,
            typename F::E* __restrict__ out) {
  typedef typename F::E E;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  E* lb = reinterpret_cast<E*>(dyn_smem);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) lb[i] = basis[i];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    const E* src = in + j * stride_j;
    typename F::Acc acc = F::acc_zero();
    uint32_t i = 0;
    while (i < n) {
      const uint32_t end = (n - i > (uint32_t)F::ACC_TERMS) ? i + F::ACC_TERMS : n;
#pragma unroll 8
      for (; i < end; ++i) F::mac(acc, src[(uint64_t)i * stride_i], lb[i]);
      if (i < n) F::acc_fold(acc);
    }
    out[j] = F::acc_reduce(acc);
  }
}

// ===================================== shamirRecoverP, Fp61, party-major, tuned
// Same contract as k_recover_p<F61> for stride_j == 1 and n <= 2048 (the C2 kernel).
// HBM-bound by design (DESIGN.md "recover kernel"):
//  * the Lagrange coefficient is split once per CTA into three 21-bit limbs
//    (b = b0 + b1*2^21 + b2*2^42) and the share into its two 32-bit words, so a
//    term is six IMAD.WIDE.U32 whose 64-bit addend IS the running sum: products are
//    < 2^53, 2048 of them fit 64 bits, no carry or reduction inside the loop;
//  * VEC = 2 secrets per thread through 128-bit loads, eight planes requested
//    before the first product is formed (the generic kernel had one load in flight);
//  * one recombination per secret: sum_k,h acc[k][h] * 2^(21k+32h), each term a
//    61-bit rotation because 2^61 = 1 (mod p).
__device__ __forceinline__ uint64_t f61_rot(uint64_t acc, int s) {  // acc * 2^s mod p, 0 <= s < 61
  uint64_t x = (acc & F61::P) + (acc >> 61);
  x = x >= F61::P ? x - F61::P : x;
  return s == 0 ? x : (((x << s) & F61::P) | (x >> (61 - s)));
}

// GatherDst (share_tc.h): where the reconstructed secrets go when the all-gather is fused into the kernel
template <int VEC>
__global__ void __launch_bounds__(256)
k_recover61_pm(const uint64_t* __restrict__ in, uint64_t N, uint32_t n, uint64_t stride_i,
               const uint64_t* __restrict__ basis, uint64_t* __restrict__ out, const __grid_constant__ GatherDst gather) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  uint4* lb = reinterpret_cast<uint4*>(dyn_smem);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
    const uint64_t b = basis[i];
    lb[i] = make_uint4((uint32_t)(b & 0x1FFFFFu), (uint32_t)((b >> 21) & 0x1FFFFFu), (uint32_t)(b >> 42), 0u);
  }
  __syncthreads();
  constexpr int BATCH = 8;
  const uint64_t units = N / VEC;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += stride) {
    uint64_t acc[VEC][6];
#pragma unroll
    for (int s = 0; s < VEC; ++s)
#pragma unroll
      for (int k = 0; k < 6; ++k) acc[s][k] = 0;
    const uint64_t* src = in + u * VEC;
    uint32_t i0 = 0;
    auto term = [&](const uint64_t (&v)[VEC], uint32_t i) {
      const uint4 b = lb[i];
#pragma unroll
      for (int s = 0; s < VEC; ++s) {
        const uint32_t a0 = (uint32_t)v[s], a1 = (uint32_t)(v[s] >> 32);
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[s][0]) : "r"(a0), "r"(b.x));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[s][1]) : "r"(a0), "r"(b.y));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[s][2]) : "r"(a0), "r"(b.z));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[s][3]) : "r"(a1), "r"(b.x));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[s][4]) : "r"(a1), "r"(b.y));
        asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[s][5]) : "r"(a1), "r"(b.z));
      }
    };
    for (; i0 + BATCH <= n; i0 += BATCH) {
      uint64_t v[BATCH][VEC];
      const uint64_t* p = src + (uint64_t)i0 * stride_i;
#pragma unroll
      for (int k = 0; k < BATCH; ++k, p += stride_i) {
        // volatile: the eight requests are issued back to back, before any product
        if constexpr (VEC == 2) {
          asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v[k][0]), "=l"(v[k][1]) : "l"(p));
        } else {
          asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v[k][0]) : "l"(p));
        }
      }
#pragma unroll
      for (int k = 0; k < BATCH; ++k) term(v[k], i0 + k);
    }
    for (; i0 < n; ++i0) {
      uint64_t v[VEC];
      const uint64_t* p = src + (uint64_t)i0 * stride_i;
#pragma unroll
      for (int s = 0; s < VEC; ++s) v[s] = __ldg(p + s);
      term(v, i0);
    }
    uint64_t r[VEC];
#pragma unroll
    for (int s = 0; s < VEC; ++s) {
      // weights 2^0, 2^21, 2^42, 2^32, 2^53, 2^74 = 2^13
      const uint64_t t = f61_rot(acc[s][0], 0) + f61_rot(acc[s][1], 21) + f61_rot(acc[s][2], 42) +
                         f61_rot(acc[s][3], 32) + f61_rot(acc[s][4], 53) + f61_rot(acc[s][5], 13);  // < 2^64
      r[s] = F61::from_raw(t);
    }
    if (gather.count == 0) {
      if constexpr (VEC == 2) {
        reinterpret_cast<ulonglong2*>(out)[u] = make_ulonglong2(r[0], r[1]);
      } else {
        out[u] = r[0];
      }
    } else {
      // all-gather fused into the reconstruction: the result goes to every rank's copy (posted stores over NVLink)
#pragma unroll
      for (uint32_t g = 0; g < 8u; ++g) {
        if (g < gather.count) {
          if constexpr (VEC == 2) {
            reinterpret_cast<ulonglong2*>(gather.dst[g])[u] = make_ulonglong2(r[0], r[1]);
          } else {
            gather.dst[g][u] = r[0];
          }
        }
      }
    }
  }
}

// ============================================================ shamirRecoverD
// shamir.h:117-140.  mat is (n_checks+1) x m: rows 0..n_checks-1 interpolate
// through shares 0..m-1 to alphas[m+r]; the last row interpolates to x.
// err[j] = 1 (and out[j] = 0) where a check row disagrees with share m+r.
template <class F>
__global__ void __launch_bounds__(256)
k_recover_d(const typename F::E* __restrict__ in, uint64_t N, uint64_t stride_i, uint64_t stride_j,
            uint32_t m, uint32_t n_checks, const typename F::E* __restrict__ mat,
            typename F::E* __restrict__ out, uint8_t* __restrict__ err,
            unsigned long long* __restrict__ n_bad) {
  typedef typename F::E E;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  E* sm = reinterpret_cast<E*>(dyn_smem);
  const uint32_t n_mat = (n_checks + 1) * m;
  for (uint32_t i = threadIdx.x; i < n_mat; i += blockDim.x) sm[i] = mat[i];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long local_bad = 0;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    const E* src = in + j * stride_j;
    bool bad = false;
    E result = F::zero();
    for (uint32_t r = 0; r <= n_checks; ++r) {
      const E* row = sm + r * m;
      typename F::Acc acc = F::acc_zero();
      uint32_t k = 0;
      while (k < m) {
        const uint32_t end = (m - k > (uint32_t)F::ACC_TERMS) ? k + F::ACC_TERMS : m;
#pragma unroll 4
        for (; k < end; ++k) F::mac(acc, src[(uint64_t)k * stride_i], row[k]);
        if (k < m) F::acc_fold(acc);
      }
      const E y = F::acc_reduce(acc);
      if (r < n_checks) {
        bad = bad || !F::eq(y, src[(uint64_t)(m + r) * stride_i]);
      } else {
        result = y;
      }
    }
    out[j] = bad ? F::zero() : result;
    err[j] = bad ? 1 : 0;
    local_bad += bad;
  }
  if (local_bad) atomicAdd(n_bad, local_bad);
}

// ============================================================ Vector entrywise
// vector.h:192-245, 522-556.  OP: 0 add, 1 subtract, 2 multiplyEntryWise
template <class F, int OP>
__global__ void __launch_bounds__(256)
k_vec_binop(const typename F::E* __restrict__ a, const typename F::E* __restrict__ b, uint64_t n,
            typename F::E* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const typename F::E x = a[i], y = b[i];
    out[i] = OP == 0 ? F::add(x, y) : OP == 1 ? F::sub(x, y) : F::mul(x, y);
  }
}

// Fp61, two elements per thread through 128-bit loads/stores (n even part)
template <int OP>
__global__ void __launch_bounds__(256)
k_vec_binop61_v2(const ulonglong2* __restrict__ a, const ulonglong2* __restrict__ b, uint64_t n2,
                 ulonglong2* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const ulonglong2 x = a[i], y = b[i];
    ulonglong2 r;
    r.x = OP == 0 ? F61::add(x.x, y.x) : OP == 1 ? F61::sub(x.x, y.x) : F61::mul(x.x, y.x);
    r.y = OP == 0 ? F61::add(x.y, y.y) : OP == 1 ? F61::sub(x.y, y.y) : F61::mul(x.y, y.y);
    out[i] = r;
  }
}

// scalarMultiply, vector.h:231-245
template <class F>
__global__ void __launch_bounds__(256)
k_vec_scale(const typename F::E* __restrict__ a, typename F::E s, uint64_t n,
            typename F::E* __restrict__ out) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = F::mul(a[i], s);
}

// z = e*b + d*a + c + e*d  (beaver.h:57-61 over Vectors)
template <class F>
__global__ void __launch_bounds__(256)
k_vec_muladd(const typename F::E* __restrict__ e, const typename F::E* __restrict__ b,
             const typename F::E* __restrict__ d, const typename F::E* __restrict__ a,
             const typename F::E* __restrict__ c, uint64_t n, typename F::E* __restrict__ z) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const typename F::E ei = e[i], di = d[i];
    typename F::Acc acc = F::acc_zero();
    F::mac(acc, ei, b[i]);
    F::mac(acc, di, a[i]);
    F::mac(acc, ei, di);
    z[i] = F::add(F::acc_reduce(acc), c[i]);
  }
}

template <>
__global__ void __launch_bounds__(256)
k_vec_muladd<F61>(const uint64_t* __restrict__ e, const uint64_t* __restrict__ b,
                  const uint64_t* __restrict__ d, const uint64_t* __restrict__ a,
                  const uint64_t* __restrict__ c, uint64_t n, uint64_t* __restrict__ z) {
  // two elements per thread, 128-bit loads/stores; odd tail by the last thread
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t n2 = n >> 1;
  const ulonglong2 *e2 = reinterpret_cast<const ulonglong2*>(e), *b2 = reinterpret_cast<const ulonglong2*>(b),
                   *d2 = reinterpret_cast<const ulonglong2*>(d), *a2 = reinterpret_cast<const ulonglong2*>(a),
                   *c2 = reinterpret_cast<const ulonglong2*>(c);
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    const ulonglong2 ev = e2[i], bv = b2[i], dv = d2[i], av = a2[i], cv = c2[i];
    F61::Acc s0 = F61::acc_zero(), s1 = F61::acc_zero();
    F61::mac(s0, ev.x, bv.x);
    F61::mac(s0, dv.x, av.x);
    F61::mac(s0, ev.x, dv.x);
    F61::mac(s1, ev.y, bv.y);
    F61::mac(s1, dv.y, av.y);
    F61::mac(s1, ev.y, dv.y);
    reinterpret_cast<ulonglong2*>(z)[i] =
        make_ulonglong2(F61::add(F61::acc_reduce(s0), cv.x), F61::add(F61::acc_reduce(s1), cv.y));
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const uint64_t i = n - 1;
    F61::Acc s = F61::acc_zero();
    F61::mac(s, e[i], b[i]);
    F61::mac(s, d[i], a[i]);
    F61::mac(s, e[i], d[i]);
    z[i] = F61::add(F61::acc_reduce(s), c[i]);
  }
}

// ------------------------------------------------------------ block reduction
template <class F>
__device__ __forceinline__ typename F::E block_reduce_sum(typename F::E v) {
  typedef typename F::E E;
  __shared__ __align__(16) unsigned char red_raw[32 * sizeof(E)];
  E* red = reinterpret_cast<E*>(red_raw);
  // warp tree through shuffles of the 64-bit words
  for (int off = 16; off > 0; off >>= 1) {
    E o;
    if constexpr (F::BYTES == 8) {
      o = __shfl_down_sync(0xffffffffu, v, off);
    } else {
      o.lo = __shfl_down_sync(0xffffffffu, v.lo, off);
      o.hi = __shfl_down_sync(0xffffffffu, v.hi, off);
    }
    v = F::add(v, o);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();  // protects red[] across repeated calls
  if (lane == 0) red[w] = v;
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    v = lane < nw ? red[lane] : F::zero();
    for (int off = 16; off > 0; off >>= 1) {
      E o;
      if constexpr (F::BYTES == 8) {
        o = __shfl_down_sync(0xffffffffu, v, off);
      } else {
        o.lo = __shfl_down_sync(0xffffffffu, v.lo, off);
        o.hi = __shfl_down_sync(0xffffffffu, v.hi, off);
      }
      v = F::add(v, o);
    }
  }
  return v;  // valid in thread 0
}

// dot (vector.h:252-259) / sum (:262-267): per-CTA partials, then a 1-CTA pass.
// b == nullptr -> sum.
template <class F>
__global__ void __launch_bounds__(256)
k_dot_partial(const typename F::E* __restrict__ a, const typename F::E* __restrict__ b, uint64_t n,
              typename F::E* __restrict__ partial) {
  typedef typename F::E E;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  E s = F::zero();
  if (b != nullptr) {
    typename F::Acc acc = F::acc_zero();
    int terms = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      F::mac(acc, a[i], b[i]);
      if (++terms == F::ACC_TERMS - 1) {
        F::acc_fold(acc);
        terms = 0;
      }
    }
    s = F::acc_reduce(acc);
  } else {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
      s = F::add(s, a[i]);
  }
  s = block_reduce_sum<F>(s);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

template <class F>
__global__ void __launch_bounds__(256)
k_sum_final(const typename F::E* __restrict__ partial, uint32_t m, typename F::E* __restrict__ out) {
  typename F::E s = F::zero();
  for (uint32_t i = threadIdx.x; i < m; i += blockDim.x) s = F::add(s, partial[i]);
  s = block_reduce_sum<F>(s);
  if (threadIdx.x == 0) out[0] = s;
}

// Vector::equals (vector.h:358-375): number of positions where a and b differ, accumulated over the
// whole vector like the reference does (no early exit); the host turns "count == 0" into the bool.
template <class F>
__global__ void __launch_bounds__(256)
k_vec_mismatch(const typename F::E* __restrict__ a, const typename F::E* __restrict__ b, uint64_t n,
               unsigned long long* __restrict__ count) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) local += !F::eq(a[i], b[i]);
  for (int off = 16; off > 0; off >>= 1) local += __shfl_down_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(count, local);
}

// ==================================================== shamirRecoverP, Fp61, SCL's own layout
// shamir.h:100-104 on a batch in SCL's [N][n] layout (row j = what shamirSecretShare returned for secret j), read
// where it lies, n <= 32.  A warp takes 32 consecutive secrets = ONE contiguous run of 32 * n elements, copies it flat
// (coalesced) into its slab of shared memory -- row stride n | 1, so that the rows can then be read by one lane each
// without bank conflicts -- and lane l forms the inner product of row l with the Lagrange basis (broadcast reads, lazy
// 128-bit accumulation: n <= 32 products).  The store is one coalesced line per warp.
// Measured (B200, 2^25 secrets, n = 32): the strided generic kernel 3.87 ms (every lane reads 16 bytes from its own
// line), lane-pair products joined by integer REDUX over 16 lanes 2.84 ms, this form 2.14 ms; the plane kernel on
// party-major input 1.43 ms.
static constexpr uint32_t kRecSmWarps = 8;

__global__ void __launch_bounds__(32 * kRecSmWarps)
k_recover61_sm(const uint64_t* __restrict__ in, uint64_t N, uint32_t n, const uint64_t* __restrict__ basis,
               uint64_t* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  uint64_t* lam = reinterpret_cast<uint64_t*>(dyn_smem);  // n coefficients, then the slabs
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  const uint32_t stride = n | 1u;
  uint64_t* slab = lam + 32 + (size_t)wid * 32u * 33u;
  if (threadIdx.x < n) lam[threadIdx.x] = basis[threadIdx.x];
  __syncthreads();
  const uint64_t warps = (uint64_t)gridDim.x * kRecSmWarps;
  const uint64_t n_runs = (N + 31u) / 32u;
  const uint32_t step_i = 32u % n, step_s = 32u / n;
  for (uint64_t q = (uint64_t)blockIdx.x * kRecSmWarps + wid; q < n_runs; q += warps) {
    const uint64_t s0 = q * 32u;
    const uint32_t cnt = (uint32_t)(N - s0 < 32u ? N - s0 : 32u);
    const uint32_t run = cnt * n;
    const uint64_t* src = in + s0 * n;
    uint32_t s = lane / n, i = lane % n;
    for (uint32_t e = lane; e < run; e += 32u) {  // flat, coalesced: element e = share i of secret s0 + s
      uint64_t v;
      asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(src + e));
      slab[s * stride + i] = v;
      i += step_i;
      s += step_s;
      if (i >= n) {
        i -= n;
        ++s;
      }
    }
    __syncwarp();
    if (lane < cnt) {
      const uint64_t* row = slab + lane * stride;
      F61::Acc acc = F61::acc_zero();
      for (uint32_t k = 0; k < n; ++k) F61::mac(acc, row[k], lam[k]);  // n <= 32 = the bound of the lazy accumulator
      out[s0 + lane] = F61::acc_reduce(acc);
    }
    __syncwarp();
  }
}

// ==================================================== Matrix::multiply(Vector)
// matrix.h:498-513: y[r] = innerProd(row r, x).  One CTA per row; A is streamed
// once with coalesced loads (8 bytes of HBM traffic per modmul), x comes from L2.
template <class F>
__global__ void __launch_bounds__(256)
k_matvec(const typename F::E* __restrict__ A, uint32_t rows, uint32_t cols,
         const typename F::E* __restrict__ x, typename F::E* __restrict__ y) {
  typedef typename F::E E;
  for (uint32_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const E* row = A + (uint64_t)r * cols;
    typename F::Acc acc = F::acc_zero();
    int terms = 0;
    for (uint32_t c = threadIdx.x; c < cols; c += blockDim.x) {
      F::mac(acc, row[c], x[c]);
      if (++terms == F::ACC_TERMS - 1) {
        F::acc_fold(acc);
        terms = 0;
      }
    }
    E s = block_reduce_sum<F>(F::acc_reduce(acc));
    if (threadIdx.x == 0) y[r] = s;
  }
}

// Fp61 form with 128-bit loads of A and x (cols even, 16-byte aligned rows)
__global__ void __launch_bounds__(256)
k_matvec61_v2(const uint64_t* __restrict__ A, uint32_t rows, uint32_t cols,
              const uint64_t* __restrict__ x, uint64_t* __restrict__ y) {
  const uint32_t c2 = cols >> 1;
  const ulonglong2* x2 = reinterpret_cast<const ulonglong2*>(x);
  for (uint32_t r = blockIdx.x; r < rows; r += gridDim.x) {
    const ulonglong2* row = reinterpret_cast<const ulonglong2*>(A + (uint64_t)r * cols);
    F61::Acc acc = F61::acc_zero();
    int terms = 0;
#pragma unroll 4
    for (uint32_t c = threadIdx.x; c < c2; c += blockDim.x) {
      const ulonglong2 av = row[c], xv = x2[c];
      F61::mac(acc, av.x, xv.x);
      F61::mac(acc, av.y, xv.y);
      terms += 2;
      if (terms >= F61::ACC_TERMS - 2) {
        F61::acc_fold(acc);
        terms = 0;
      }
    }
    uint64_t s = block_reduce_sum<F61>(F61::acc_reduce(acc));
    if (threadIdx.x == 0) y[r] = s;
  }
}

// Fp61, one WARP per row (cols even, 16-byte aligned rows): four independent 128-bit loads
// of A in flight per lane, x through the read-only path (64 KiB at C5: L1/L2 resident), a
// shuffle tree at the end -- no block barrier between rows.  HBM-bound: 8 bytes per modmul.
__global__ void __launch_bounds__(256)
k_matvec61_warp(const uint64_t* __restrict__ A, uint32_t rows, uint32_t cols,
                const uint64_t* __restrict__ x, uint64_t* __restrict__ y) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t c2 = cols >> 1;
  const ulonglong2* x2 = reinterpret_cast<const ulonglong2*>(x);
  for (uint32_t r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += warps) {
    const ulonglong2* row = reinterpret_cast<const ulonglong2*>(A + (uint64_t)r * cols);
    F61::Acc acc = F61::acc_zero();
    uint32_t c = lane;
    for (; c + 96u < c2; c += 128u) {
      ulonglong2 a[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(a[k].x), "=l"(a[k].y) : "l"(row + c + 32 * k));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const ulonglong2 xv = __ldg(x2 + c + 32 * k);
        F61::mac(acc, a[k].x, xv.x);
        F61::mac(acc, a[k].y, xv.y);
      }
      F61::acc_fold(acc);  // 8 products per round: far below the 32-term bound
    }
    for (; c < c2; c += 32u) {
      const ulonglong2 av = __ldg(row + c), xv = __ldg(x2 + c);
      F61::mac(acc, av.x, xv.x);
      F61::mac(acc, av.y, xv.y);
      F61::acc_fold(acc);
    }
    uint64_t s = F61::acc_reduce(acc);
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s = F61::add(s, __shfl_down_sync(0xffffffffu, s, off));
    if (lane == 0) y[r] = s;
  }
}

// Fp61 mat-vec in two passes for long rows (cols a multiple of 512, 16-byte aligned rows).  Pass 1: the matrix is cut
// into 4 KiB chunks in MEMORY order (a chunk lies inside one row) and warp w takes chunks w, w + W, ...: the grid
// sweeps A front to back like the vector kernels do, eight 128-bit loads in flight per lane, no barrier, and writes one
// partial sum per chunk.  Pass 2: one thread per row adds the row's partials (cols / 512 of them, contiguous).
// Measured alternatives at 8192 x 8192 (B200): one warp per row 0.107 ms (every concurrent access in another row),
// two warps per row joined through shared memory 0.116 ms, a 512-thread CTA per row 0.144 ms.
static constexpr uint32_t kMatvecChunk = 256;  // 16-byte units per chunk: 4 KiB = 512 elements

__global__ void __launch_bounds__(256)
k_matvec61_chunks(const uint64_t* __restrict__ A, uint64_t n_chunks, uint32_t chunks_per_row,
                  const uint64_t* __restrict__ x, uint64_t* __restrict__ partial) {
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const ulonglong2* A2 = reinterpret_cast<const ulonglong2*>(A);
  const ulonglong2* x2 = reinterpret_cast<const ulonglong2*>(x);
  for (uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < n_chunks; q += warps) {
    const ulonglong2* src = A2 + q * kMatvecChunk + lane;
    const ulonglong2* xs = x2 + (q % chunks_per_row) * kMatvecChunk + lane;
    ulonglong2 a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(a[k].x), "=l"(a[k].y) : "l"(src + 32 * k));
    F61::Acc acc = F61::acc_zero();
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const ulonglong2 xv = __ldg(xs + 32 * k);
      F61::mac(acc, a[k].x, xv.x);
      F61::mac(acc, a[k].y, xv.y);
    }
    uint64_t s = F61::acc_reduce(acc);  // 16 products: below the 32-term bound of the lazy accumulator
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s = F61::add(s, __shfl_down_sync(0xffffffffu, s, off));
    if (lane == 0) partial[q] = s;
  }
}

// The same sweep with ROUNDS x 4 KiB per chunk (a chunk still lies inside one row: cols a multiple of 512 * ROUNDS), the
// column offset carried from chunk to chunk instead of a 64-bit modulo per chunk, and the 32 lane sums joined by three
// warp-wide integer reductions (REDUX) over 21-bit limbs instead of five shuffle + modular-add levels: about a fifth
// fewer instructions per byte than k_matvec61_chunks.  MINB: resident CTAs per SM the register budget is set for --
// with 4 (64 registers) the loads of a chunk really are in flight together; at 32 registers ptxas serialises them.
// Measured at 8192 x 8192, L2 flushed before every call, one call per event pair (tools/matvec_probe.py):
// k_matvec61_chunks 0.109 ms, this kernel at 32 registers 0.105, at 64 registers 0.0965 (16 loads in flight at 120
// registers: the same).
template <int ROUNDS, int MINB>
__global__ void __launch_bounds__(256, MINB)
k_matvec61_sweep(const uint64_t* __restrict__ A, uint64_t n_chunks, uint32_t chunks_per_row,
                 const uint64_t* __restrict__ x, uint64_t* __restrict__ partial) {
  constexpr uint32_t CH = kMatvecChunk * ROUNDS;  // 16-byte units per chunk
  const uint32_t lane = threadIdx.x & 31u;
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const ulonglong2* A2 = reinterpret_cast<const ulonglong2*>(A);
  const ulonglong2* x2 = reinterpret_cast<const ulonglong2*>(x);
  uint64_t q = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  uint32_t xoff = (uint32_t)(q % chunks_per_row);
  const uint32_t step = (uint32_t)(warps % chunks_per_row);
  for (; q < n_chunks; q += warps) {
    const ulonglong2* src = A2 + q * CH + lane;
    const ulonglong2* xs = x2 + (uint64_t)xoff * CH + lane;
    F61::Acc acc = F61::acc_zero();  // 16 * ROUNDS <= 32 products: inside the bound of the lazy accumulator
    ulonglong2 a[8 * ROUNDS];  // every load of the chunk is requested before the first product
#pragma unroll
    for (int k = 0; k < 8 * ROUNDS; ++k)
      asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(a[k].x), "=l"(a[k].y) : "l"(src + 32 * k));
#pragma unroll
    for (int k = 0; k < 8 * ROUNDS; ++k) {
      const ulonglong2 xv = __ldg(xs + 32 * k);
      F61::mac(acc, a[k].x, xv.x);
      F61::mac(acc, a[k].y, xv.y);
    }
    const uint64_t s = F61::acc_reduce(acc);  // < 2^61
    const uint32_t t0 = __reduce_add_sync(0xffffffffu, (uint32_t)s & 0x1FFFFFu);          // 32 * 2^21 = 2^26
    const uint32_t t1 = __reduce_add_sync(0xffffffffu, (uint32_t)(s >> 21) & 0x1FFFFFu);
    const uint32_t t2 = __reduce_add_sync(0xffffffffu, (uint32_t)(s >> 42));              // 32 * 2^19 = 2^24
    if (lane == 0)  // t2 * 2^42 = (t2 mod 2^19) * 2^42 + (t2 >> 19) * 2^61, and 2^61 = 1: the sum stays below 2^62
      partial[q] = F61::from_raw((uint64_t)t0 + ((uint64_t)t1 << 21) + ((uint64_t)(t2 & 0x7FFFFu) << 42) + (uint64_t)(t2 >> 19));
    xoff += step;
    if (xoff >= chunks_per_row) xoff -= chunks_per_row;
  }
}

__global__ void __launch_bounds__(256)
k_matvec61_finish(const uint64_t* __restrict__ partial, uint32_t rows, uint32_t chunks_per_row, uint64_t* __restrict__ y) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const uint64_t* p = partial + (uint64_t)r * chunks_per_row;
  uint64_t s = 0;
  for (uint32_t k = 0; k < chunks_per_row; ++k) s = F61::add(s, p[k]);
  y[r] = s;
}

// Matrix::vandermonde(n, m), xs = 1..n (matrix.h:445-460): thread = one row
template <class F>
__global__ void k_vandermonde(uint32_t n, uint32_t m, typename F::E* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const typename F::E x = F::from_u32(i + 1);
  typename F::E v = F::one();
  for (uint32_t j = 0; j < m; ++j) {
    out[(uint64_t)i * m + j] = v;
    v = F::mul(v, x);
  }
}

// Matrix::vandermonde(n, m, xs) (matrix.h:445-460): row i = (1, xs[i], xs[i]^2, ...); thread = one row
template <class F>
__global__ void k_vandermonde_xs(uint32_t n, uint32_t m, const typename F::E* __restrict__ xs,
                                 typename F::E* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const typename F::E x = xs[i];
  typename F::E v = F::one();
  for (uint32_t j = 0; j < m; ++j) {
    out[(uint64_t)i * m + j] = v;
    v = F::mul(v, x);
  }
}

// Polynomial::evaluate (poly.h:56-64) of N polynomials at n caller-chosen points: Horner from the top coefficient.
// coeffs: planes [t+1][N] (coefficient k of polynomial j at k*N + j); value (j, i) at out[i*stride_i + j*stride_j].
// The points sit in shared memory; thread = one polynomial, its coefficients re-read per point (L1/L2 resident).
template <class F>
__global__ void __launch_bounds__(256)
k_poly_eval(const typename F::E* __restrict__ coeffs, uint64_t N, uint32_t t, const typename F::E* __restrict__ xs,
            uint32_t n, typename F::E* __restrict__ out, uint64_t stride_i, uint64_t stride_j) {
  typedef typename F::E E;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  E* sx = reinterpret_cast<E*>(dyn_smem);
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) sx[i] = xs[i];
  __syncthreads();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; j < N; j += stride) {
    E* dst = out + j * stride_j;
    for (uint32_t i = 0; i < n; ++i) {
      const E x = sx[i];
      E r = coeffs[(uint64_t)t * N + j];
      for (int64_t k = (int64_t)t - 1; k >= 0; --k) r = F::add(coeffs[(uint64_t)k * N + j], F::mul(r, x));
      dst[(uint64_t)i * stride_i] = r;
    }
  }
}

// ================================================================== transpose
// [rows][cols] -> [cols][rows], 32x32 tiles through padded shared memory
template <class E>
__global__ void __launch_bounds__(256)
k_transpose(const E* __restrict__ in, uint64_t rows, uint64_t cols, E* __restrict__ out) {
  __shared__ E tile[32][33];
  const uint64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  const uint64_t n_tiles = tiles_c * tiles_r;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (uint64_t tid = blockIdx.x; tid < n_tiles; tid += gridDim.x) {
    const uint64_t tr = tid / tiles_c, tc = tid % tiles_c;
    for (uint32_t k = ty; k < 32; k += 8) {
      const uint64_t r = tr * 32 + k, c = tc * 32 + tx;
      if (r < rows && c < cols) tile[k][tx] = in[r * cols + c];
    }
    __syncthreads();
    for (uint32_t k = ty; k < 32; k += 8) {
      const uint64_t c = tc * 32 + k, r = tr * 32 + tx;
      if (r < rows && c < cols) out[c * rows + r] = tile[tx][k];
    }
    __syncthreads();
  }
}

// The same transposition when one side is NARROW (a batch of sharings: [n][N] planes <-> SCL's [N][n], n <= 32).  The
// 32 x 32 tiles above keep 32 - n lanes idle on the narrow side (Fp61: 4.8 TB/s at n = 32, 2.9 at 16, 1.3 at 5).  Here
// a WARP owns 32 consecutive positions of the long side; the other side of its data is ONE contiguous run of 32 * n
// elements, so it is moved flat -- lane l takes elements l, l + 32, ... of the run -- with the (position, row) index
// carried from step to step instead of divided out.  A private slab of shared memory per warp, __syncwarp only.
// TO_LONG_ROWS: in = [long][n] (runs), out = [n][long] (planes); otherwise the reverse.
static constexpr uint32_t kNarrowWarps = 4;    // per CTA: 4 slabs of 32 x 33 elements = 33 KiB for Fp61, 66 KiB for Fp127

template <class E, bool TO_LONG_ROWS>
__global__ void __launch_bounds__(32 * kNarrowWarps)
k_transpose_narrow(const E* __restrict__ in, uint64_t n_long, uint32_t n, E* __restrict__ out) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  E* slab = reinterpret_cast<E*>(dyn_smem) + (size_t)wid * 32u * 33u;  // [row i < n][position s < 32], row stride 33
  const uint64_t warps = (uint64_t)gridDim.x * kNarrowWarps;
  const uint64_t groups = (n_long + 31u) / 32u;
  const uint32_t step_i = 32u % n, step_s = 32u / n;
  for (uint64_t g = (uint64_t)blockIdx.x * kNarrowWarps + wid; g < groups; g += warps) {
    const uint64_t s0 = g * 32u;
    const uint32_t cnt = (uint32_t)(n_long - s0 < 32u ? n_long - s0 : 32u);  // positions of this group
    const uint32_t run = cnt * n;
    if (TO_LONG_ROWS) {
      const E* src = in + s0 * n;
      uint32_t s = lane / n, i = lane % n;
      for (uint32_t e = lane; e < run; e += 32u) {  // flat, coalesced
        slab[i * 33u + s] = src[e];
        i += step_i;
        s += step_s;
        if (i >= n) {
          i -= n;
          ++s;
        }
      }
      __syncwarp();
      if (lane < cnt)
        for (uint32_t r = 0; r < n; ++r) out[(uint64_t)r * n_long + s0 + lane] = slab[r * 33u + lane];
    } else {
      if (lane < cnt)
        for (uint32_t r = 0; r < n; ++r) slab[r * 33u + lane] = in[(uint64_t)r * n_long + s0 + lane];
      __syncwarp();
      E* dst = out + s0 * n;
      uint32_t s = lane / n, i = lane % n;
      for (uint32_t e = lane; e < run; e += 32u) {
        dst[e] = slab[i * 33u + s];
        i += step_i;
        s += step_s;
        if (i >= n) {
          i -= n;
          ++s;
        }
      }
    }
    __syncwarp();
  }
}

// [rows][cols][W] -> [cols][rows][W]: the same transposition on W-wide elements
// (Vector<Array<FF, W>> shares, secret-major <-> party-major).  32 x 32 tiles of
// up to CW components at a time, contiguous runs of 32*cw elements on both sides.
template <class E, int CW>
__global__ void __launch_bounds__(256)
k_transpose_wide(const E* __restrict__ in, uint64_t rows, uint64_t cols, uint32_t W, E* __restrict__ out) {
  __shared__ E tile[32][32 * CW + 1];
  const uint64_t tiles_c = (cols + 31) / 32, tiles_r = (rows + 31) / 32;
  const uint32_t n_chunks = (W + CW - 1) / CW;
  const uint64_t n_tiles = tiles_c * tiles_r * n_chunks;
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (uint64_t tid = blockIdx.x; tid < n_tiles; tid += gridDim.x) {
    const uint32_t w0 = (uint32_t)(tid % n_chunks) * CW;
    const uint32_t cw = (W - w0 < (uint32_t)CW) ? W - w0 : (uint32_t)CW;
    const uint64_t tc = (tid / n_chunks) % tiles_c, tr = tid / (n_chunks * tiles_c);
    for (uint32_t k = wid; k < 32; k += 8) {
      const uint64_t r = tr * 32 + k;
      if (r >= rows) break;
      for (uint32_t e = lane; e < 32 * cw; e += 32) {
        const uint32_t q = e / cw, w = e % cw;
        const uint64_t c = tc * 32 + q;
        if (c < cols) tile[k][q * CW + w] = in[(r * cols + c) * W + w0 + w];
      }
    }
    __syncthreads();
    for (uint32_t k = wid; k < 32; k += 8) {
      const uint64_t c = tc * 32 + k;
      if (c >= cols) break;
      for (uint32_t e = lane; e < 32 * cw; e += 32) {
        const uint32_t q = e / cw, w = e % cw;
        const uint64_t r = tr * 32 + q;
        if (r < rows) out[(c * rows + r) * W + w0 + w] = tile[q][k * CW + w];
      }
    }
    __syncthreads();
  }
}

// ================================================= integer-pipe microbenchmark
// kind 0: IMAD (32-bit), 1: IMAD.WIDE.U32, 2: LOP3, 3: IADD3, 4: LDS.32 (conflict-free)
template <int KIND>
__global__ void __launch_bounds__(512, 1) k_pipe_bench(uint32_t iters, uint32_t* sink) {
  __shared__ uint32_t sm[1024];
  uint32_t a0 = threadIdx.x + 1, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13,
           a6 = a0 * 17, a7 = a0 * 19;
  const uint32_t m = blockIdx.x | 1;
  const uint32_t lane_sm = smem_u32(sm) + (threadIdx.x & 31) * 4;
  if (KIND == 4) {
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 37 + 32) & 0xFE0;
    __syncthreads();
  }
  if (KIND == 1) {
    uint64_t w0 = a0, w1 = a1, w2 = a2, w3 = a3, w4 = a4, w5 = a5, w6 = a6, w7 = a7;
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w0) : "r"((uint32_t)w0), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w1) : "r"((uint32_t)w1), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w2) : "r"((uint32_t)w2), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w3) : "r"((uint32_t)w3), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w4) : "r"((uint32_t)w4), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w5) : "r"((uint32_t)w5), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w6) : "r"((uint32_t)w6), "r"(m));
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w7) : "r"((uint32_t)w7), "r"(m));
      }
    }
    a0 = (uint32_t)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) ^ (uint32_t)((w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) >> 32);
  } else {
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
#define SCLGPU_PB(v_)                                                                        \
  if (KIND == 0) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(v_) : "r"(m));            \
  if (KIND == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v_) : "r"(m), "r"(it)); \
  if (KIND == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(v_) : "r"(m));                   \
  if (KIND == 4) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v_) : "r"(lane_sm + (v_ & 0xF80)));
        SCLGPU_PB(a0) SCLGPU_PB(a1) SCLGPU_PB(a2) SCLGPU_PB(a3) SCLGPU_PB(a4) SCLGPU_PB(a5)
        SCLGPU_PB(a6) SCLGPU_PB(a7)
#undef SCLGPU_PB
      }
    }
  }
  const uint32_t r = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
  if (r == 0x12345678u) sink[0] = r;  // practically never: keeps the chains alive
}

}  // namespace sclgpu
