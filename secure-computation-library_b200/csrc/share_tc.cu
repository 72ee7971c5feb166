// share_tc.cu -- the tcgen05 (5th-generation tensor core) kernels of the Shamir path:
//   k_share_tcm<F, ...>      shamirSecretShare, both fields, PRG fused or coefficient planes (the default)
//   k_share61_tc             first variant, A operand staged in shared memory (kept as a measured comparison)
//   k_recover_d_tc<F, ...>   shamirRecoverD, and shamirRecoverP for Fp127
//
// Reference path replaced: ss::shamirSecretShare (include/scl/ss/shamir.h:52-68) =
// Vector::random(t+1, prg) (vector.h:508-519, prg.cc:124-146), c[0] = secret,
// n Horner evaluations at x = 1..n (poly.h:56-64) -- N calls on one PRG; and
// ss::shamirRecoverD / shamirRecoverP (shamir.h:82-155).
//
// Formulation.  The shares of one secret are the Vandermonde
// product  share_i = sum_k c_k * (i+1)^k  (the reference's own
// test/scl/math/test_matrix.cc:342-365 states this identity).  A coefficient is
// used through its BYTES c_k = sum_a c_{k,a} 2^(8a) -- the raw little-endian
// keystream word, no reduction needed because the map is linear -- and the constant
//     C_{i,k,a} = (i+1)^k * 2^(8a) mod p
// through its bytes C_{i,k,a} = sum_s C_{i,k,a,s} 2^(8s).  Then
//     share_i = sum_s 2^(8s) * acc_{i,s},   acc_{i,s} = sum_{k,a} c_{k,a} * C_{i,k,a,s}
// and acc is a u8 x u8 -> s32 matrix product with K = BYTES*(t+1) <= 128 (acc < 2^23):
//     D[128 secrets][parties x limbs] += A[128 secrets][K] * B[K][64 per pass]
// issued as tcgen05.mma.kind::i8 (M=128, N=64, K=32 per instruction), A = the
// coefficient bytes exactly as the PRG emits them (one 128-byte row per secret),
// B = the constant limbs (32 KiB, K-major, 128B-swizzled, built on the host once per
// (field, t, n), resident in shared memory), D in tensor memory.  The epilogue reads the
// 23-bit limbs of a share with tcgen05.ld and recombines them mod p.
//
// Default schedule (k_share_tcm<F, 5, 1, 64>): one CTA per SM, 640 threads = 5 groups
// of 4 warps; a group owns one 128-secret tile at a time: every thread draws its
// secret's keystream (T-table AES-CTR, aes_ctr.cuh) and writes it straight into its own
// TENSOR-MEMORY lane (tcgen05.st; the A operand never touches shared memory), one
// elected thread issues the `.ts` MMAs for the group, each warp drains its own 32 TMEM
// lanes.  Groups run unsynchronised with respect to each other (named barriers and one
// mbarrier per group), so one group's AES (LSU + ALU pipes) overlaps another's MMA and
// epilogue.  The reconstruction kernels use the same structure with the A rows loaded
// from the share planes instead of drawn from the PRG.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "aes_ctr.cuh"
#include "field.cuh"
#include "share_tc.h"
#include "tc_common.cuh"

namespace sclgpu {

static constexpr uint32_t kTcATile = 16384;  // 128 rows x 128 B

__global__ void __launch_bounds__(kTcThreads, 1)
k_share61_tc(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0,
             const uint4* __restrict__ g_bmat, uint64_t first_block, const uint64_t* __restrict__ secrets,
             uint64_t N, uint32_t t, uint32_t n, uint64_t* __restrict__ out, uint64_t stride_i,
             uint64_t stride_j) {
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t dyn = smem_u32(dyn_smem);
  const uint32_t tbase = aes_table_base(dyn_smem);        // AES tables, 64 KiB aligned (aes_ctr.cuh)
  const uint32_t a_base = (dyn + 1023u) & ~1023u;         // 3 A tiles below the tables
  const uint32_t b_base = tbase + kAesTableBytes;         // B limbs above them
  const uint32_t ctl = b_base + kTcBmatBytes;             // 6 mbarriers + the TMEM base address
  if (a_base + kTcGroups * kTcATile > tbase || ctl + 64u > dyn + kTcDynSmem) __trap();

  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  aes_fill_tables(tbase, g_t0);
  for (uint32_t e = tid; e < kTcBmatBytes / 16; e += kTcThreads) {
    const uint4 w = __ldg(g_bmat + e);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(b_base + e * 16u), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ctl + 48u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (uint32_t i = 0; i < 2 * kTcGroups; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ctl + 8u * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // B was written through the generic proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(ctl + 48u) : "memory");

  uint32_t lanebase = tbase + (tid & 31u) * 4u;
  asm volatile("" : "+r"(lanebase)::"memory");

  const uint32_t g = tid >> 7, gt = tid & 127u;            // group, thread in group = row of the tile
  const uint32_t a_tile = a_base + g * kTcATile;
  const uint32_t xr = gt & 7u;
  const uint32_t row = a_tile + (gt >> 3) * 1024u + xr * 128u;
  const uint32_t acc0 = tmem + g * (2u * kTcPassCols);     // two 64-column accumulators
  const uint32_t lane_off = ((warp & 3u) * 32u) << 16;     // this warp's TMEM lanes
  const uint32_t mbar0 = ctl + 16u * g, mbar1 = mbar0 + 8u;
  uint32_t ph0 = 0, ph1 = 0;

  const uint32_t nblk = ((t + 1u) * 8u + 15u) / 16u;       // keystream blocks per secret (prg.cc:129-133)
  const uint32_t ksteps = ((t + 1u) * 8u + 31u) / 32u;     // K = 32 bytes per MMA
  const uint32_t npass = (n + 7u) / 8u;
  const uint64_t tiles = (N + 127u) / 128u;

  auto issue_pass = [&](uint32_t p) {  // one elected thread
    const uint32_t d = acc0 + (p & 1u) * kTcPassCols;
    for (uint32_t ks = 0; ks < ksteps; ++ks)
      tc_mma(d, tc_desc(a_tile + ks * 32u), tc_desc(b_base + p * (kTcPassCols * 128u) + ks * 32u), ks);
    tc_commit((p & 1u) ? mbar1 : mbar0);
  };

  for (uint64_t tile = (uint64_t)blockIdx.x * kTcGroups + g; tile < tiles; tile += (uint64_t)gridDim.x * kTcGroups) {
    const uint64_t j = tile * 128u + gt;
    const bool valid = j < N;
    if (valid) {
      const uint64_t sec = secrets[j];
      const uint64_t ctr0 = first_block + j * nblk;
      if (t == 0) {
        asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %3};" ::"r"(row + (xr << 4)), "r"((uint32_t)sec), "r"((uint32_t)(sec >> 32)), "r"(0u) : "memory");
      } else {
        PrgGroup grp;
        uint64_t gid = ctr0 >> 8;
        prg_group(key, lanebase, ctr0, grp);
#pragma unroll 1
        for (uint32_t b = 0; b < nblk; ++b) {
          const uint64_t ctr = ctr0 + b;
          if ((ctr >> 8) != gid) {  // crossed a 256-block group: at most once per secret
            gid = ctr >> 8;
            prg_group(key, lanebase, ctr, grp);
          }
          uint32_t o0, o1, o2, o3;
          prg_block_grouped(key, lanebase, grp, (uint32_t)ctr, o0, o1, o2, o3);
          if (b == 0) {  // slot 0 of the draw is replaced by the secret (shamir.h:56-57)
            o0 = (uint32_t)sec;
            o1 = (uint32_t)(sec >> 32);
          }
          // block b = coefficients 2b, 2b+1 = 16-byte chunk b of the row (chunk ^ row%8: 128B swizzle)
          asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(row + ((b ^ xr) << 4)), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        }
      }
    }
    __syncwarp();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // A rows -> visible to the tensor core
    group_sync(g);
    if (gt == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_pass(0);
      if (npass > 1) issue_pass(1);
    }
    for (uint32_t p = 0; p < npass; ++p) {
      if (p & 1u) {
        mbar_wait(mbar1, ph1);
        ph1 ^= 1u;
      } else {
        mbar_wait(mbar0, ph0);
        ph0 ^= 1u;
      }
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[64];
      tmem_ld64(acc0 + (p & 1u) * kTcPassCols + lane_off, v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      group_sync(g);  // every warp of the group has drained this accumulator
      if (gt == 0 && p + 2u < npass) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_pass(p + 2u);
      }
      if (valid) {
        uint64_t* dst = out + j * stride_j + (uint64_t)(p * 8u) * stride_i;
#pragma unroll
        for (uint32_t ii = 0; ii < 8; ++ii) {
          if (p * 8u + ii < n) dst[(uint64_t)ii * stride_i] = tc_combine(v + 8 * ii);
        }
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}


// ---------------------------------------------------------------------------
// Variant with the A operand in TENSOR MEMORY: every thread writes its secret's
// coefficient bytes into its own TMEM lane (tcgen05.st, 4 columns per keystream
// block), and the MMA reads A from TMEM (`tcgen05.mma ... [d], [a], b_desc`).
// Shared memory then carries only the table lookups and the B limbs: the A-tile
// stores and the four A re-reads per tile of the variant above (5 of its 41 LSU
// wavefronts per secret) disappear, and with no A tiles in shared memory a fourth
// group of warps fits.  TMEM per group: 32 columns of A + NBUF accumulators of 64.
// GROUPS x 128 threads; NBUF accumulators of PCOLS columns (PCOLS / F::BYTES parties per MMA pass) per
// group.  F = F61: K = 8(t+1) bytes, t <= 15, n <= 32.  F = F127: K = 16(t+1) bytes, t <= 7, n <= 16;
// coefficient k is keystream block k of the secret's draw (vector.h:508-519 with 16-byte elements).
// COEFFS = false: coefficients drawn from the PRG (shamirSecretShare).  COEFFS = true: coefficient planes
// supplied by the caller (Polynomial::create + evaluate, poly.h:56-64,179-198): `secrets` is then the
// [t+1][N] array (coefficient k of secret j at k*N + j), no AES tables, and the kernel is HBM-bound.
// MODE 2 (array-valued secrets, shamirSecretShare on math::Array<FF, W>; pedersen.h:137-138): the N "secrets"
// are the N/W sharings' components, v = j*W + c; sharing j draws (t+1)*W stream elements, element k*W + c being
// coefficient k of component c (array.h:82-88 inside vector.h:508-519).  Fp127: one block per element, so thread v
// draws blocks ctr0(j) + k*W + c itself.  Fp61 (W even): block (k*W + c)/2 holds coefficient k of components c
// and c^1, which sit in adjacent lanes: the even lane draws the blocks of coefficients 0..h-1, the odd lane those
// of h..2h-1 (h = ceil((t+1)/2)), and the two exchange the halves they do not own with one shuffle pair per block
// -- the same AES work per thread as the plain kernel.
template <class F, int GROUPS, int NBUF, int PCOLS, int MODE>
__global__ void __launch_bounds__(128 * GROUPS, 1)
k_share_tcm(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0,
            const uint4* __restrict__ g_bmat, uint64_t first_block, const typename F::E* __restrict__ secrets,
            uint64_t N, uint32_t t, uint32_t n, typename F::E* __restrict__ out, uint64_t stride_i,
            uint64_t stride_j, uint32_t W) {
  typedef typename F::E E;
  constexpr bool COEFFS = MODE == 1, WIDE = MODE == 2;
  constexpr uint32_t EB = F::BYTES;                        // bytes = 8-bit limbs per element
  constexpr uint32_t kThreads = 128 * GROUPS;
  constexpr uint32_t kColsPerGroup = 32u + NBUF * PCOLS;
  constexpr uint32_t kPassParties = PCOLS / EB;
  constexpr uint32_t kLdParties = 32u / EB;                // parties per 32-column TMEM load
  static_assert(GROUPS * kColsPerGroup <= 512, "tensor memory has 512 columns");
  static_assert(PCOLS == 32 || PCOLS == 64, "pass width");
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t dyn = smem_u32(dyn_smem);
  const uint32_t tbase = COEFFS ? ((dyn + 1023u) & ~1023u) : aes_table_base(dyn_smem);
  const uint32_t b_base = COEFFS ? tbase : tbase + kAesTableBytes;
  const uint32_t ctl = b_base + kTcBmatBytes;             // 2*GROUPS mbarriers + the TMEM base address
  if (ctl + 128u > dyn + (COEFFS ? kTcCoeffDynSmem : kTcmDynSmem)) __trap();

  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  if (!COEFFS) aes_fill_tables(tbase, g_t0);
  for (uint32_t e = tid; e < kTcBmatBytes / 16; e += kThreads) {
    const uint4 w = __ldg(g_bmat + e);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(b_base + e * 16u), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ctl + 120u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (uint32_t i = 0; i < 2 * GROUPS; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ctl + 8u * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(ctl + 120u) : "memory");

  uint32_t lanebase = tbase + (tid & 31u) * 4u;
  asm volatile("" : "+r"(lanebase)::"memory");

  const uint32_t g = tid >> 7, gt = tid & 127u;
  const uint32_t a_tm = tmem + g * kColsPerGroup;          // 32 columns: 128 coefficient bytes per lane
  const uint32_t acc0 = a_tm + 32u;
  const uint32_t lane_off = ((warp & 3u) * 32u) << 16;
  const uint32_t mbar0 = ctl + 16u * g, mbar1 = mbar0 + 8u;
  uint32_t ph0 = 0, ph1 = 0;

  const uint32_t nblk = ((t + 1u) * EB + 15u) / 16u;       // keystream blocks per secret (prg.cc:129-133)
  const uint32_t ksteps = ((t + 1u) * EB + 31u) / 32u;     // K = 32 bytes per MMA
  const uint32_t npass = (n + kPassParties - 1u) / kPassParties;
  const uint64_t tiles = (N + 127u) / 128u;

  auto issue_pass = [&](uint32_t p) {
    const uint32_t b = (NBUF == 2) ? (p & 1u) : 0u;
    for (uint32_t ks = 0; ks < ksteps; ++ks)
      tc_mma_ts(acc0 + b * PCOLS, a_tm + ks * 8u, tc_desc(b_base + p * (PCOLS * 128u) + ks * 32u), tc_idesc(PCOLS), ks);
    tc_commit(b ? mbar1 : mbar0);
  };
  // SCL's own layout as the output (stride_i == 1: the shares of a secret are adjacent): 16-byte stores of two shares
  // -- a warp's store then touches 32 rows, and half as many of them as 8-byte stores would
  const bool pair_stores = EB == 8 && stride_i == 1 && (stride_j & 1u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  auto emit = [&](const uint32_t (&v)[32], E* dst, uint32_t first_party) {
    if constexpr (EB == 8) {
      if (pair_stores && first_party + kLdParties <= n) {
#pragma unroll
        for (uint32_t ii = 0; ii < kLdParties; ii += 2)
          *reinterpret_cast<ulonglong2*>(dst + ii) = make_ulonglong2(tc_combine(v + 8 * ii), tc_combine(v + 8 * ii + 8));
        return;
      }
    }
    if (first_party + kLdParties <= n) {  // all parties of this load exist (always, when n is a multiple of 4 / 2)
#pragma unroll
      for (uint32_t ii = 0; ii < kLdParties; ++ii) {
        if constexpr (EB == 8) {
          dst[(uint64_t)ii * stride_i] = tc_combine(v + 8 * ii);
        } else {
          dst[(uint64_t)ii * stride_i] = tc_combine127(v + 16 * ii);
        }
      }
      return;
    }
#pragma unroll
    for (uint32_t ii = 0; ii < kLdParties; ++ii) {
      if (first_party + ii < n) {
        if constexpr (EB == 8) {
          dst[(uint64_t)ii * stride_i] = tc_combine(v + 8 * ii);
        } else {
          dst[(uint64_t)ii * stride_i] = tc_combine127(v + 16 * ii);
        }
      }
    }
  };

  // COEFFS: the (t+1) coalesced plane loads of a tile are requested one tile ahead (right after the previous tile's
  // rows went to tensor memory), so they are in flight under that tile's MMAs and epilogue
  constexpr uint32_t kMaxC = COEFFS ? 128u / EB : 1u;
  E c[kMaxC];
  auto load_coeffs = [&](uint64_t tile_) {
    const uint64_t j_ = tile_ * 128u + gt;
    const E* p = secrets + (j_ < N ? j_ : N - 1);
#pragma unroll
    for (uint32_t k = 0; k < kMaxC; ++k, p += N) c[k] = (k <= t) ? *p : F::zero();
  };
  const uint64_t tile_first = (uint64_t)blockIdx.x * GROUPS + g, tile_step = (uint64_t)gridDim.x * GROUPS;
  if constexpr (COEFFS) {
    if (tile_first < tiles) load_coeffs(tile_first);
  }

  for (uint64_t tile = tile_first; tile < tiles; tile += tile_step) {
    const uint64_t j = tile * 128u + gt;
    const bool valid = j < N;
    const uint64_t jj = valid ? j : N - 1;                 // tail lanes recompute the last secret (never stored)
    const uint32_t a_lane = a_tm + lane_off;
    if constexpr (COEFFS) {
      // 16 bytes per tcgen05.st
      if constexpr (EB == 8) {
#pragma unroll
        for (uint32_t q = 0; q < kMaxC / 2; ++q)
          if (2 * q <= t)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane + 4u * q),
                         "r"((uint32_t)c[2 * q]), "r"((uint32_t)(c[2 * q] >> 32)), "r"((uint32_t)c[2 * q + 1]),
                         "r"((uint32_t)(c[2 * q + 1] >> 32))
                         : "memory");
      } else {
#pragma unroll
        for (uint32_t q = 0; q < kMaxC; ++q)
          if (q <= t)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane + 4u * q),
                         "r"((uint32_t)c[q].lo), "r"((uint32_t)(c[q].lo >> 32)), "r"((uint32_t)c[q].hi),
                         "r"((uint32_t)(c[q].hi >> 32))
                         : "memory");
      }
    } else {
    uint32_t s0, s1, s2 = 0, s3 = 0;                       // the secret = coefficient 0 (shamir.h:56-57)
    if constexpr (EB == 8) {
      const uint64_t sec = secrets[jj];
      s0 = (uint32_t)sec;
      s1 = (uint32_t)(sec >> 32);
    } else {
      const E sec = secrets[jj];
      s0 = (uint32_t)sec.lo;
      s1 = (uint32_t)(sec.lo >> 32);
      s2 = (uint32_t)sec.hi;
      s3 = (uint32_t)(sec.hi >> 32);
    }
    if constexpr (WIDE) {
      const uint64_t sh = jj / W;                          // the sharing, and the component within it
      const uint32_t c = (uint32_t)(jj - sh * W);
      PrgGroup grp;
      if constexpr (EB == 16) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane), "r"(s0), "r"(s1), "r"(s2), "r"(s3) : "memory");
        const uint64_t ctr0 = first_block + sh * ((uint64_t)(t + 1u) * W) + c;
        uint64_t gid = ~0ull;
#pragma unroll 1
        for (uint32_t k = 1; k <= t; ++k) {
          const uint64_t ctr = ctr0 + (uint64_t)k * W;
          if ((ctr >> 8) != gid) {
            gid = ctr >> 8;
            prg_group(key, lanebase, ctr, grp);
          }
          uint32_t o0, o1, o2, o3;
          prg_block_grouped(key, lanebase, grp, (uint32_t)ctr, o0, o1, o2, o3);
          __syncwarp();
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane + 4u * k), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        }
      } else {
        const uint32_t h = (t + 2u) / 2u, odd = c & 1u, hw = W / 2u;
        // block of coefficient k of this component pair: base + k*(W/2)
        const uint64_t base = first_block + sh * ((uint64_t)(t + 1u) * hw) + (c >> 1);
        uint64_t gid = ~0ull;
#pragma unroll 1
        for (uint32_t i = 0; i < h; ++i) {
          const uint64_t ctr = base + (uint64_t)(odd ? h + i : i) * hw;  // k = h + i > t on the odd lane: drawn, unused
          if ((ctr >> 8) != gid) {
            gid = ctr >> 8;
            prg_group(key, lanebase, ctr, grp);
          }
          uint32_t o0, o1, o2, o3;
          prg_block_grouped(key, lanebase, grp, (uint32_t)ctr, o0, o1, o2, o3);
          __syncwarp();
          // words 0,1 = the even component, words 2,3 = the odd one: keep mine, hand the others to the neighbour
          const uint32_t r0 = __shfl_xor_sync(0xFFFFFFFFu, odd ? o0 : o2, 1);
          const uint32_t r1 = __shfl_xor_sync(0xFFFFFFFFu, odd ? o1 : o3, 1);
          uint32_t lo0 = odd ? r0 : o0, lo1 = odd ? r1 : o1;   // coefficient i
          const uint32_t hi0 = odd ? o2 : r0, hi1 = odd ? o3 : r1;  // coefficient h + i
          if (i == 0) {
            lo0 = s0;
            lo1 = s1;
          }
          asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a_lane + 2u * i), "r"(lo0), "r"(lo1) : "memory");
          asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a_lane + 2u * (h + i)), "r"(hi0), "r"(hi1) : "memory");
        }
      }
    } else {
    const uint64_t ctr0 = first_block + jj * nblk;
    if (t == 0 || EB == 16) {
      // Fp61, t = 0: one block is consumed, none of it is used.  Fp127: block 0 is slot 0 of the draw,
      // consumed and replaced by the secret.
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane), "r"(s0), "r"(s1), "r"(s2), "r"(s3) : "memory");
    }
    if (t != 0) {
      PrgGroup grp;
      const uint32_t b0 = (EB == 16) ? 1u : 0u;
      prg_group(key, lanebase, ctr0 + b0, grp);
      // the block of this secret (if any; nblk <= 8) that opens the next 256-counter group
      const uint32_t cross = b0 + 256u - ((uint32_t)(ctr0 + b0) & 255u), c_lo = (uint32_t)ctr0;
      // Blocks go two at a time when no secret of the warp crosses a 256-counter group (the usual case): two independent
      // lookup chains per thread for the scheduler to interleave -- with one block at a time a warp stalls at every round
      // boundary (XOR tree -> address -> lookup).  An odd count (Fp127: blocks 1..7) starts with one single block.
      const bool pairs = __all_sync(0xffffffffu, cross >= nblk);
      const uint32_t single_end = pairs ? b0 + ((nblk - b0) & 1u) : nblk;
      uint32_t b = b0;
#pragma unroll 1
      for (; b < single_end; ++b) {
        if (b == cross) prg_group(key, lanebase, ctr0 + b, grp);  // at most once per secret
        uint32_t o0 = s0, o1 = s1, o2, o3;  // Fp61 block 0: coefficient 0 is the secret, the keystream words are not computed
        prg_block_grouped(key, lanebase, grp, c_lo + b, o0, o1, o2, o3, !(EB == 8 && b == 0));
        __syncwarp();
        // keystream block b = K bytes [16b, 16b+16) of the row = TMEM columns 4b..4b+3 of this lane
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane + 4u * b), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
      }
#pragma unroll 1
      for (; b < nblk; b += 2u) {
        uint32_t o0 = s0, o1 = s1, o2, o3, q0, q1, q2, q3;
        prg_block_grouped(key, lanebase, grp, c_lo + b, o0, o1, o2, o3, !(EB == 8 && b == 0));
        prg_block_grouped(key, lanebase, grp, c_lo + b + 1u, q0, q1, q2, q3);
        __syncwarp();
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a_lane + 4u * b),
                     "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(q0), "r"(q1), "r"(q2), "r"(q3) : "memory");
      }
    }
    }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    group_sync(g);
    if (gt == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_pass(0);
      if (NBUF == 2 && npass > 1) issue_pass(1);
    }
    if constexpr (COEFFS) {
      if (tile + tile_step < tiles) load_coeffs(tile + tile_step);
    }
    for (uint32_t p = 0; p < npass; ++p) {
      if (NBUF == 2 && (p & 1u)) {
        mbar_wait(mbar1, ph1);
        ph1 ^= 1u;
      } else {
        mbar_wait(mbar0, ph0);
        ph0 ^= 1u;
      }
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = acc0 + ((NBUF == 2) ? (p & 1u) : 0u) * PCOLS + lane_off;
      E* dst = out + j * stride_j + (uint64_t)(p * kPassParties) * stride_i;
      uint32_t v[32];
      tmem_ld32(acc, v);
      if (PCOLS == 32) {
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync(g);  // accumulator drained by every warp of the group: it may be overwritten
        if (gt == 0 && p + NBUF < npass) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_pass(p + NBUF);
        }
      }
      if (valid) emit(v, dst, p * kPassParties);
      if (PCOLS == 64) {
        tmem_ld32(acc + 32u, v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync(g);
        if (gt == 0 && p + NBUF < npass) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_pass(p + NBUF);
        }
        if (valid) emit(v, dst + (uint64_t)kLdParties * stride_i, p * kPassParties + kLdParties);
      }
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---------------------------------------------------------------------------
// Warp-specialised variant (variant 4).  The groups of k_share_tcm alternate between the AES phase (LSU + ALU)
// and the MMA / epilogue phase, so the LSU -- the limiting pipe -- idles whenever several groups happen to be in
// their epilogues.  Here the roles are fixed: 4 PRODUCER groups (16 warps) only draw keystream into double-
// buffered A rows in tensor memory, 2 CONSUMER groups (8 warps) only issue the MMAs, drain the accumulators and
// store; consumer c serves producers 2c and 2c+1.  Hand-off through mbarriers: fullA[pg][buf] (128 producer
// arrivals after tcgen05.wait::st), emptyA[pg][buf] (tcgen05.commit after the tile's last MMA), dfull[c][b]
// (tcgen05.commit per pass); the consumer's own accumulator reuse is ordered by a named barrier.
// TMEM: 4 x 2 x 32 columns of A + 2 x 2 x 64 of D = 512.  Measured (2^26, n=32, t=15): 11.67 ms against 11.06 ms for
// k_share_tcm<F61,5,1,64> -- the unspecialised groups already keep ~16 warps in the AES phase, and the limit is the
// joint saturation of the LSU and ALU pipes, not phase bubbles.  Kept selectable (SCLGPU_SHARE_TC=4) and tested.
static constexpr int kWsProducers = 4, kWsConsumers = 2;
static constexpr int kWsThreads = 128 * (kWsProducers + kWsConsumers);

template <class F>
__global__ void __launch_bounds__(kWsThreads, 1)
k_share_ws(const __grid_constant__ AesKey key, const uint32_t* __restrict__ g_t0, const uint4* __restrict__ g_bmat,
           uint64_t first_block, const typename F::E* __restrict__ secrets, uint64_t N, uint32_t t, uint32_t n,
           typename F::E* __restrict__ out, uint64_t stride_i, uint64_t stride_j) {
  typedef typename F::E E;
  constexpr uint32_t EB = F::BYTES;
  constexpr uint32_t PCOLS = 64, kPassParties = PCOLS / EB, kLdParties = 32u / EB;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t dyn = smem_u32(dyn_smem);
  const uint32_t tbase = aes_table_base(dyn_smem);
  const uint32_t b_base = tbase + kAesTableBytes;
  const uint32_t ctl = b_base + kTcBmatBytes;  // fullA[4][2] | emptyA[4][2] | dfull[2][2] | TMEM address
  if (ctl + 256u > dyn + kTcmDynSmem + 128u) __trap();
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  const uint32_t bar_full = ctl, bar_empty = ctl + 64u, bar_dfull = ctl + 128u, tmem_slot = ctl + 160u;

  aes_fill_tables(tbase, g_t0);
  for (uint32_t e = tid; e < kTcBmatBytes / 16; e += kWsThreads) {
    const uint4 w = __ldg(g_bmat + e);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(b_base + e * 16u), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (uint32_t i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(bar_full + 8u * i) : "memory");
    for (uint32_t i = 0; i < 8; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_empty + 8u * i) : "memory");
    for (uint32_t i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_dfull + 8u * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

  const uint32_t nblk = ((t + 1u) * EB + 15u) / 16u;
  const uint32_t ksteps = ((t + 1u) * EB + 31u) / 32u;
  const uint32_t npass = (n + kPassParties - 1u) / kPassParties;
  const uint64_t tiles = (N + 127u) / 128u;
  const uint64_t tile_stride = (uint64_t)gridDim.x * kWsProducers;
  const uint32_t gt = tid & 127u;
  const uint32_t lane_off = ((warp & 3u) * 32u) << 16;

  if (warp < 4u * kWsProducers) {
    // ------------------------------------------------------------------ producer: AES-CTR -> A rows in TMEM
    uint32_t lanebase = tbase + (tid & 31u) * 4u;
    asm volatile("" : "+r"(lanebase)::"memory");
    const uint32_t pg = warp >> 2;
    uint32_t it = 0;
    for (uint64_t tile = (uint64_t)blockIdx.x * kWsProducers + pg; tile < tiles; tile += tile_stride, ++it) {
      const uint32_t buf = it & 1u;
      if (it >= 2u) mbar_wait(bar_empty + 8u * (pg * 2u + buf), ((it >> 1) - 1u) & 1u);  // MMAs of tile it-2 are done with this buffer
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t j = tile * 128u + gt;
      const uint64_t jj = j < N ? j : N - 1;
      const uint32_t a_lane = tmem + pg * 64u + buf * 32u + lane_off;
      uint32_t s0, s1, s2 = 0, s3 = 0;
      if constexpr (EB == 8) {
        const uint64_t sec = secrets[jj];
        s0 = (uint32_t)sec;
        s1 = (uint32_t)(sec >> 32);
      } else {
        const E sec = secrets[jj];
        s0 = (uint32_t)sec.lo;
        s1 = (uint32_t)(sec.lo >> 32);
        s2 = (uint32_t)sec.hi;
        s3 = (uint32_t)(sec.hi >> 32);
      }
      const uint64_t ctr0 = first_block + jj * nblk;
      if (t == 0 || EB == 16)
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane), "r"(s0), "r"(s1), "r"(s2), "r"(s3) : "memory");
      if (t != 0) {
        PrgGroup grp;
        const uint32_t b0 = (EB == 16) ? 1u : 0u;
        uint64_t gid = (ctr0 + b0) >> 8;
        prg_group(key, lanebase, ctr0 + b0, grp);
#pragma unroll 1
        for (uint32_t b = b0; b < nblk; ++b) {
          const uint64_t ctr = ctr0 + b;
          if ((ctr >> 8) != gid) {
            gid = ctr >> 8;
            prg_group(key, lanebase, ctr, grp);
          }
          uint32_t o0, o1, o2, o3;
          prg_block_grouped(key, lanebase, grp, (uint32_t)ctr, o0, o1, o2, o3);
          if (EB == 8 && b == 0) {
            o0 = s0;
            o1 = s1;
          }
          __syncwarp();
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane + 4u * b), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        }
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_full + 8u * (pg * 2u + buf)) : "memory");
    }
  } else {
    // ------------------------------------------------------------------ consumer: MMA issue, drain, recombine, store
    const uint32_t cg = (warp >> 2) - kWsProducers;       // 0, 1
    const uint32_t d0 = tmem + 256u + cg * 128u;          // two 64-column accumulators
    const uint32_t df0 = bar_dfull + 16u * cg, df1 = df0 + 8u;
    uint32_t ph0 = 0, ph1 = 0;
    auto emit = [&](const uint32_t (&v)[32], E* dst, uint32_t first_party) {
#pragma unroll
      for (uint32_t ii = 0; ii < kLdParties; ++ii) {
        if (first_party + ii < n) {
          if constexpr (EB == 8) dst[(uint64_t)ii * stride_i] = tc_combine(v + 8 * ii);
          else dst[(uint64_t)ii * stride_i] = tc_combine127(v + 16 * ii);
        }
      }
    };
    // tiles in the order the two producers make them: (it, pg) with pg = 2cg, 2cg+1
    for (uint32_t it = 0;; ++it) {
      bool any = false;
      for (uint32_t q = 0; q < 2u; ++q) {
        const uint32_t pg = 2u * cg + q;
        const uint64_t tile = (uint64_t)blockIdx.x * kWsProducers + pg + (uint64_t)it * tile_stride;
        if (tile >= tiles) continue;
        any = true;
        const uint32_t buf = it & 1u;
        const uint32_t a_tm = tmem + pg * 64u + buf * 32u;
        auto issue_pass = [&](uint32_t p) {
          const uint32_t b = p & 1u;
          for (uint32_t ks = 0; ks < ksteps; ++ks)
            tc_mma_ts(d0 + b * PCOLS, a_tm + ks * 8u, tc_desc(b_base + p * (PCOLS * 128u) + ks * 32u), tc_idesc(PCOLS), ks);
          tc_commit(b ? df1 : df0);
          if (p + 1u == npass) tc_commit(bar_empty + 8u * (pg * 2u + buf));  // every MMA reading this A buffer is issued
        };
        mbar_wait(bar_full + 8u * (pg * 2u + buf), (it >> 1) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (gt == 0) {
          issue_pass(0);
          if (npass > 1) issue_pass(1);
        }
        const uint64_t j = tile * 128u + gt;
        const bool valid = j < N;
        for (uint32_t p = 0; p < npass; ++p) {
          if (p & 1u) {
            mbar_wait(df1, ph1);
            ph1 ^= 1u;
          } else {
            mbar_wait(df0, ph0);
            ph0 ^= 1u;
          }
          __syncwarp();
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t acc = d0 + (p & 1u) * PCOLS + lane_off;
          E* dst = out + j * stride_j + (uint64_t)(p * kPassParties) * stride_i;
          uint32_t v[32];
          tmem_ld32(acc, v);
          if (valid) emit(v, dst, p * kPassParties);
          tmem_ld32(acc + 32u, v);
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1u + cg) : "memory");  // accumulator drained by all four consumer warps
          if (gt == 0 && p + 2u < npass) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_pass(p + 2u);
          }
          if (valid) emit(v, dst + (uint64_t)kLdParties * stride_i, p * kPassParties + kLdParties);
        }
      }
      if (!any) break;
    }
  }

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// (variant, groups, accumulators, columns per pass).  Measured at 2^26 secrets, n=32, t=15 on B200
// (gpurun, CUDA events): A in shared memory 3 groups 13.45 ms | (3,2,64) 12.72 | (4,1,64) 11.70 |
// (4,2,32) 13.11 | (5,1,64) 11.33 | (5,2,32) 12.42 | (6,1,32) 12.27 | (7,1,32) 12.25 | (8,1,32) 12.19.
// Kept: the two best.
#define SCLGPU_TCM_VARIANTS(X) X(2, 4, 1, 64) X(3, 5, 1, 64)

// ===================================================== shamirRecoverD on tensor cores
// shamir.h:117-140: y_r = sum_{k<m} shares[k] * L_r[k] for the n_checks rows interpolating to
// alphas[m+r] (compared with shares[m+r]) and the final row interpolating to x.  The same
// byte-limb product as the share kernel: A = the first m shares of a secret (m*BYTES <= 128
// bytes, written by its thread into its TMEM lane), B = bytes of L_r[k] * 2^(8a) (built on the
// host from the device-computed Lagrange rows), D = (n_checks+1) x BYTES limb columns.  No AES,
// so shared memory holds only B; the kernel is HBM-bound (reads d+t shares per secret once).
// Measured and dropped (round 2): the next tile's m shares requested right after the stores of the current one (as the
// reconstruction group of k_share_recover61 does) costs m * BYTES / 4 more registers, i.e. four groups instead of five:
// Fp127 n = 16 t = 7, 2^24 secrets 0.882 ms against 0.822 ms for this form.
template <class F, int GROUPS, int ACOLS>
__global__ void __launch_bounds__(128 * GROUPS, 1)
k_recover_d_tc(const uint4* __restrict__ g_bmat, const typename F::E* __restrict__ in, uint64_t N,
               uint64_t stride_i, uint64_t stride_j, uint32_t m, uint32_t n_checks,
               typename F::E* __restrict__ out, uint8_t* __restrict__ err,
               unsigned long long* __restrict__ n_bad) {
  typedef typename F::E E;
  constexpr uint32_t EB = F::BYTES;
  constexpr uint32_t kThreads = 128 * GROUPS;
  constexpr uint32_t PCOLS = 64;
  constexpr uint32_t kACols = ACOLS;                       // 32: A rows of 128 bytes; 64: up to 256 bytes = two K tiles
  constexpr uint32_t kColsPerGroup = kACols + PCOLS;
  constexpr uint32_t kMaxM = 4u * ACOLS / EB;              // shares in one A row
  constexpr uint32_t kLdRows = 32u / EB;                   // output rows per 32-column TMEM load
  static_assert(GROUPS * kColsPerGroup <= 512, "tensor memory has 512 columns");
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t dyn = smem_u32(dyn_smem);
  const uint32_t b_base = (dyn + 1023u) & ~1023u;
  const uint32_t ctl = b_base + kTcBmatBytes;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t e = tid; e < kTcBmatBytes / 16; e += kThreads) {
    const uint4 w = __ldg(g_bmat + e);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(b_base + e * 16u), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ctl + 120u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (uint32_t i = 0; i < GROUPS; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ctl + 8u * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(ctl + 120u) : "memory");

  const uint32_t g = tid >> 7, gt = tid & 127u;
  const uint32_t a_tm = tmem + g * kColsPerGroup;
  const uint32_t acc0 = a_tm + kACols;
  const uint32_t lane_off = ((warp & 3u) * 32u) << 16;
  const uint32_t a_lane = a_tm + lane_off;
  const uint32_t mbar = ctl + 8u * g;
  uint32_t ph = 0;
  const uint32_t ksteps = (m * EB + 31u) / 32u;
  const uint32_t rows = n_checks + 1u;
  const uint32_t npass = (rows * EB + PCOLS - 1u) / PCOLS;
  const uint64_t tiles = (N + 127u) / 128u;
  unsigned long long local_bad = 0;

  auto issue_pass = [&](uint32_t p) {
    for (uint32_t ks = 0; ks < ksteps; ++ks)
      tc_mma_ts(acc0, a_tm + ks * 8u, tc_desc(b_base + (ks >> 2) * kTcRdKTileBytes + p * (PCOLS * 128u) + (ks & 3u) * 32u),
                tc_idesc(PCOLS), ks);
    tc_commit(mbar);
  };

  const uint32_t n_planes = m + n_checks;
  for (uint64_t tile = (uint64_t)blockIdx.x * GROUPS + g; tile < tiles; tile += (uint64_t)gridDim.x * GROUPS) {
    const uint64_t j = tile * 128u + gt;
    const bool valid = j < N;
    const E* src = in + (valid ? j : N - 1) * stride_j;
    if (stride_j == 1) {
      // party-major planes: the group's NEXT tile is pulled into L2 now (128 * BYTES bytes per plane = BYTES lines of
      // 128 bytes), so that its loads, a whole tile of MMAs and epilogue later, are L2 hits -- no registers held
      const uint64_t jn = (tile + (uint64_t)gridDim.x * GROUPS) * 128u;
      for (uint32_t line = gt; line < n_planes * EB; line += 128u) {
        const uint64_t e = jn + (uint64_t)(line % EB) * (128u / EB);
        if (e < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(in + (uint64_t)(line / EB) * stride_i + e));
      }
    }
    // the m interpolation shares: all requested before the first is consumed
    E a[kMaxM];
#pragma unroll
    for (uint32_t k = 0; k < kMaxM; ++k) a[k] = (k < m) ? src[(uint64_t)k * stride_i] : F::zero();
    if constexpr (EB == 8) {
#pragma unroll
      for (uint32_t c = 0; c < kMaxM / 2; ++c) {
        if (2 * c < m)  // warp-uniform
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane + 4u * c),
                       "r"((uint32_t)a[2 * c]), "r"((uint32_t)(a[2 * c] >> 32)), "r"((uint32_t)a[2 * c + 1]),
                       "r"((uint32_t)(a[2 * c + 1] >> 32))
                       : "memory");
      }
    } else {
#pragma unroll
      for (uint32_t c = 0; c < kMaxM; ++c) {
        if (c < m)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane + 4u * c),
                       "r"((uint32_t)a[c].lo), "r"((uint32_t)(a[c].lo >> 32)), "r"((uint32_t)a[c].hi),
                       "r"((uint32_t)(a[c].hi >> 32))
                       : "memory");
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    group_sync(g);
    if (gt == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      issue_pass(0);
    }
    bool bad = false;
    E result = F::zero();
    for (uint32_t p = 0; p < npass; ++p) {
      // the shares this pass is checked against: requested before waiting for the MMA
      constexpr uint32_t kPassRows = PCOLS / EB;
      E chk[kPassRows];
#pragma unroll
      for (uint32_t rr = 0; rr < kPassRows; ++rr) {
        const uint32_t r = p * kPassRows + rr;
        chk[rr] = (r < n_checks) ? src[(uint64_t)(m + r) * stride_i] : F::zero();
      }
      mbar_wait(mbar, ph);
      ph ^= 1u;
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[32];
#pragma unroll
      for (uint32_t h = 0; h < 2; ++h) {
        tmem_ld32(acc0 + lane_off + 32u * h, v);
        if (h == 1) {
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          group_sync(g);  // accumulator drained: the next pass may overwrite it
          if (gt == 0 && p + 1u < npass) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            issue_pass(p + 1u);
          }
        }
#pragma unroll
        for (uint32_t ii = 0; ii < kLdRows; ++ii) {
          const uint32_t rr = h * kLdRows + ii, r = p * kPassRows + rr;
          E y;
          if constexpr (EB == 8) y = (ACOLS == 64) ? tc_combine24(v + 8 * ii) : tc_combine(v + 8 * ii);  // 24-bit limbs with two K tiles
          else y = tc_combine127(v + 16 * ii);
          if (r < n_checks) bad = bad || !F::eq(y, chk[rr]);
          else if (r == n_checks) result = y;
        }
      }
    }
    if (valid) {
      out[j] = bad ? F::zero() : result;
      if (err != nullptr) err[j] = bad ? 1 : 0;
      local_bad += bad;
    }
  }
  if (local_bad && n_bad != nullptr) atomicAdd(n_bad, local_bad);

  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

static constexpr int kRdGroups = 4;
static constexpr uint32_t kRdDynSmem = 120u * 1024u;  // > half of an SM's shared memory: one CTA (one TMEM owner) per SM

template <class F>
static cudaError_t recover_d_tc_launch_t(cudaStream_t st, int sm_count, const void* d_bmat, const typename F::E* d_in,
                                         uint64_t N, uint64_t si, uint64_t sj, uint32_t m, uint32_t n_checks,
                                         typename F::E* d_out, uint8_t* d_err, unsigned long long* d_count) {
  const uint64_t tiles = (N + 127) / 128;
  const uint4* bm = reinterpret_cast<const uint4*>(d_bmat);
  // per launch: the attribute belongs to the current device's instance of the kernel, and this is a few microseconds
  if (m * F::BYTES <= 128u) {  // one K tile: 32 columns of A, five groups
    constexpr int G = 5;
    cudaError_t e = cudaFuncSetAttribute(k_recover_d_tc<F, G, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRdDynSmem);
    if (e != cudaSuccess) return e;
    const int grid = (int)std::min<uint64_t>((tiles + G - 1) / G, (uint64_t)sm_count);
    k_recover_d_tc<F, G, 32><<<grid, 128 * G, kRdDynSmem, st>>>(bm, d_in, N, si, sj, m, n_checks, d_out, d_err, d_count);
  } else {
    cudaError_t e = cudaFuncSetAttribute(k_recover_d_tc<F, kRdGroups, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRdDynSmem);
    if (e != cudaSuccess) return e;
    const int grid = (int)std::min<uint64_t>((tiles + kRdGroups - 1) / kRdGroups, (uint64_t)sm_count);
    k_recover_d_tc<F, kRdGroups, 64><<<grid, 128 * kRdGroups, kRdDynSmem, st>>>(bm, d_in, N, si, sj, m, n_checks, d_out, d_err, d_count);
  }
  return cudaGetLastError();
}

cudaError_t recover_d61_tc_launch(cudaStream_t st, int sm_count, const void* d_bmat, const uint64_t* d_in, uint64_t N,
                                  uint64_t si, uint64_t sj, uint32_t m, uint32_t n_checks, uint64_t* d_out, uint8_t* d_err,
                                  unsigned long long* d_count) {
  return recover_d_tc_launch_t<F61>(st, sm_count, d_bmat, d_in, N, si, sj, m, n_checks, d_out, d_err, d_count);
}
cudaError_t recover_d127_tc_launch(cudaStream_t st, int sm_count, const void* d_bmat, const E127* d_in, uint64_t N,
                                   uint64_t si, uint64_t sj, uint32_t m, uint32_t n_checks, E127* d_out, uint8_t* d_err,
                                   unsigned long long* d_count) {
  return recover_d_tc_launch_t<F127>(st, sm_count, d_bmat, d_in, N, si, sj, m, n_checks, d_out, d_err, d_count);
}

cudaError_t share_tc_prepare() {
  cudaError_t e = cudaFuncSetAttribute(k_share61_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcDynSmem);
#define X(V, G, NB, PC)                                                                                                                       \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_share_tcm<F61, G, NB, PC, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcmDynSmem); \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_share_tcm<F127, G, NB, PC, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcmDynSmem);
  SCLGPU_TCM_VARIANTS(X)
#undef X
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_share_ws<F61>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTcmDynSmem + 128u));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_share_ws<F127>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kTcmDynSmem + 128u));
  return e;
}

int tc_variant_groups(int variant) {
  if (variant == 4) return kWsProducers;
#define X(V, G, NB, PC) \
  if (variant == V) return G;
  SCLGPU_TCM_VARIANTS(X)
#undef X
  return kTcGroups;
}

cudaError_t share61_tc_launch(int variant, cudaStream_t st, int grid, const AesKey& key, const uint32_t* d_t0, const void* d_bmat,
                              uint64_t first_block, const uint64_t* d_secrets, uint64_t N, uint32_t t, uint32_t n,
                              uint64_t* d_out, uint64_t stride_i, uint64_t stride_j) {
  const uint4* bm = reinterpret_cast<const uint4*>(d_bmat);
  if (variant == 4) {
    k_share_ws<F61><<<grid, kWsThreads, kTcmDynSmem + 128u, st>>>(key, d_t0, bm, first_block, d_secrets, N, t, n, d_out, stride_i, stride_j);
    return cudaGetLastError();
  }
  bool done = false;
#define X(V, G, NB, PC)                                                                                               \
  if (variant == V) {                                                                                                 \
    k_share_tcm<F61, G, NB, PC, 0><<<grid, 128 * G, kTcmDynSmem, st>>>(key, d_t0, bm, first_block, d_secrets, N, t, n,     \
                                                                  d_out, stride_i, stride_j, 1u);                     \
    done = true;                                                                                                      \
  }
  SCLGPU_TCM_VARIANTS(X)
#undef X
  if (!done)
    k_share61_tc<<<grid, kTcThreads, kTcDynSmem, st>>>(key, d_t0, bm, first_block, d_secrets, N, t, n, d_out, stride_i, stride_j);
  return cudaGetLastError();
}

cudaError_t share127_tc_launch(int variant, cudaStream_t st, int grid, const AesKey& key, const uint32_t* d_t0,
                               const void* d_bmat, uint64_t first_block, const E127* d_secrets, uint64_t N, uint32_t t,
                               uint32_t n, E127* d_out, uint64_t stride_i, uint64_t stride_j) {
  const uint4* bm = reinterpret_cast<const uint4*>(d_bmat);
  if (variant == 4) {
    k_share_ws<F127><<<grid, kWsThreads, kTcmDynSmem + 128u, st>>>(key, d_t0, bm, first_block, d_secrets, N, t, n, d_out, stride_i, stride_j);
    return cudaGetLastError();
  }
  if (variant == 2) {
    k_share_tcm<F127, 4, 1, 64, 0><<<grid, 512, kTcmDynSmem, st>>>(key, d_t0, bm, first_block, d_secrets, N, t, n, d_out, stride_i, stride_j, 1u);
  } else {
    k_share_tcm<F127, 5, 1, 64, 0><<<grid, 640, kTcmDynSmem, st>>>(key, d_t0, bm, first_block, d_secrets, N, t, n, d_out, stride_i, stride_j, 1u);
  }
  return cudaGetLastError();
}

// Polynomial evaluation from caller-supplied coefficient planes on the tensor cores (no PRG)
template <class F>
static cudaError_t share_coeffs_tc_launch_t(cudaStream_t st, int sm_count, const void* d_bmat, const typename F::E* d_coeffs,
                                            uint64_t N, uint32_t t, uint32_t n, typename F::E* d_out, uint64_t stride_i,
                                            uint64_t stride_j) {
  auto kern = k_share_tcm<F, 4, 1, 64, 1>;  // four groups: the prefetched coefficients stay in registers
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcCoeffDynSmem);
  if (e != cudaSuccess) return e;
  const uint64_t tiles = (N + 127) / 128;
  const int grid = (int)std::min<uint64_t>((tiles + 3) / 4, (uint64_t)sm_count);
  AesKey unused{};
  kern<<<grid, 512, kTcCoeffDynSmem, st>>>(unused, nullptr, reinterpret_cast<const uint4*>(d_bmat), 0, d_coeffs, N, t, n, d_out,
                                          stride_i, stride_j, 1u);
  return cudaGetLastError();
}
cudaError_t share61_coeffs_tc_launch(cudaStream_t st, int sm_count, const void* d_bmat, const uint64_t* d_coeffs, uint64_t N,
                                     uint32_t t, uint32_t n, uint64_t* d_out, uint64_t stride_i, uint64_t stride_j) {
  return share_coeffs_tc_launch_t<F61>(st, sm_count, d_bmat, d_coeffs, N, t, n, d_out, stride_i, stride_j);
}
cudaError_t share127_coeffs_tc_launch(cudaStream_t st, int sm_count, const void* d_bmat, const E127* d_coeffs, uint64_t N,
                                      uint32_t t, uint32_t n, E127* d_out, uint64_t stride_i, uint64_t stride_j) {
  return share_coeffs_tc_launch_t<F127>(st, sm_count, d_bmat, d_coeffs, N, t, n, d_out, stride_i, stride_j);
}

// shamirSecretShare on math::Array<FF, W>: NW = N*W component polynomials, PRG fused (k_share_tcm MODE 2).
// Fp61 needs an even W (the neighbour exchange); Fp127 takes any W.
template <class F>
static cudaError_t share_wide_tc_launch_t(cudaStream_t st, int sm_count, const AesKey& key, const uint32_t* d_t0, const void* d_bmat,
                                          uint64_t first_block, const typename F::E* d_secrets, uint64_t NW, uint32_t W,
                                          uint32_t t, uint32_t n, typename F::E* d_out, uint64_t stride_i, uint64_t stride_j) {
  auto kern = k_share_tcm<F, 5, 1, 64, 2>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcmDynSmem);
  if (e != cudaSuccess) return e;
  const uint64_t tiles = (NW + 127) / 128;
  const int grid = (int)std::min<uint64_t>((tiles + 4) / 5, (uint64_t)sm_count);
  kern<<<grid, 640, kTcmDynSmem, st>>>(key, d_t0, reinterpret_cast<const uint4*>(d_bmat), first_block, d_secrets, NW, t, n, d_out,
                                      stride_i, stride_j, W);
  return cudaGetLastError();
}
cudaError_t share61_wide_tc_launch(cudaStream_t st, int sm_count, const AesKey& key, const uint32_t* d_t0, const void* d_bmat,
                                   uint64_t first_block, const uint64_t* d_secrets, uint64_t NW, uint32_t W, uint32_t t, uint32_t n,
                                   uint64_t* d_out, uint64_t stride_i, uint64_t stride_j) {
  return share_wide_tc_launch_t<F61>(st, sm_count, key, d_t0, d_bmat, first_block, d_secrets, NW, W, t, n, d_out, stride_i, stride_j);
}
cudaError_t share127_wide_tc_launch(cudaStream_t st, int sm_count, const AesKey& key, const uint32_t* d_t0, const void* d_bmat,
                                    uint64_t first_block, const E127* d_secrets, uint64_t NW, uint32_t W, uint32_t t, uint32_t n,
                                    E127* d_out, uint64_t stride_i, uint64_t stride_j) {
  return share_wide_tc_launch_t<F127>(st, sm_count, key, d_t0, d_bmat, first_block, d_secrets, NW, W, t, n, d_out, stride_i, stride_j);
}

}  // namespace sclgpu
