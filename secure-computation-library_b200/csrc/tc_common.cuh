// tc_common.cuh -- device helpers shared by the tcgen05 kernels of the Shamir path (share_tc.cu,
// share_recover.cu): shared-memory / instruction descriptors of tcgen05.mma.kind::i8, the MMA / commit /
// mbarrier wrappers, tensor-memory loads, and the recombination of the 23-bit limb accumulators mod p.
#pragma once
#include <cstdint>

#include "field.cuh"

namespace sclgpu {

static constexpr uint32_t kTcPassCols = 64;  // 8 parties x 8 limbs per MMA pass

// K-major, SWIZZLE_128B shared-memory matrix descriptor: start address >> 4,
// LBO = 1 (unused for swizzled K-major), SBO = 1024 B (8 rows), version 1 (sm_100).
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}

// instruction descriptor: D = s32, A = B = u8, both K-major, M = 128, N = 64
static constexpr uint32_t tc_idesc(uint32_t n_cols) { return (2u << 4) | ((n_cols >> 3) << 17) | ((128u >> 4) << 24); }
static constexpr uint32_t kTcIdesc = tc_idesc(kTcPassCols);

__device__ __forceinline__ void tc_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kTcIdesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tc_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}

// The suspend-time hint lets the hardware park the warp until the phase completes (or the hint expires) instead of
// returning after the default, much shorter, limit: the share kernels executed ~18 try_wait round trips per wait
// (ncu: 37.7 M SYNCS for 2.1 M waits), issue slots the AES warps of the same SM sub-partition want.
#ifndef SCLGPU_MBAR_HINT_NS
#define SCLGPU_MBAR_HINT_NS 20000
#endif
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mbar), "r"(parity), "r"((uint32_t)SCLGPU_MBAR_HINT_NS)
        : "memory");
  } while (!ok);
}

__device__ __forceinline__ void group_sync(uint32_t g) {
  asm volatile("bar.sync %0, 128;" ::"r"(g + 1u) : "memory");
}

// 32 lanes x 64 columns of 32 bits: thread l of the warp gets lane (base + l)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]),
        "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]),
        "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]),
        "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]),
        "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
        "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
        "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
        "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// share = sum_s v[s] * 2^(8s) mod p, v[s] < 2^23; canonical result.
// R = the limbs gathered below 2^64 (2^61 = 1 folds the top of limb pair 6,7); then with a = R >> 61 (<= 5)
// q = floor(R / p) = (R + a + 1) >> 61 exactly, and R mod p = (R + q) mod 2^61 -- no compare / select.
__device__ __forceinline__ uint64_t tc_combine(const uint32_t* v) {
  const uint32_t p01 = v[0] + (v[1] << 8), p23 = v[2] + (v[3] << 8);  // < 2^32
  const uint32_t p45 = v[4] + (v[5] << 8), p67 = v[6] + (v[7] << 8);
  // p67 * 2^48 = (p67 mod 2^13) * 2^48 + (p67 >> 13) * 2^61, and 2^61 = 1
  const uint64_t R = (uint64_t)p01 + ((uint64_t)p23 << 16) + ((uint64_t)p45 << 32) +
                     ((uint64_t)(p67 & 0x1FFFu) << 48) + (uint64_t)(p67 >> 13);  // < 2^63 + 2^62
  const uint32_t lo = (uint32_t)R, hi = (uint32_t)(R >> 32);
  const uint32_t a1 = (hi >> 29) + 1u;
  uint32_t t0, q, r0, r1;
  asm("{\n\t"
      "add.cc.u32 %0, %4, %6;\n\t"        // R + a + 1: only the bits from 61 up are used
      "addc.u32 %1, %5, 0;\n\t"
      "shr.u32 %1, %1, 29;\n\t"           // q
      "add.cc.u32 %2, %4, %1;\n\t"        // R + q
      "addc.u32 %3, %5, 0;\n\t"
      "and.b32 %3, %3, 0x1FFFFFFF;\n\t"
      "}"
      : "=&r"(t0), "=&r"(q), "=&r"(r0), "=&r"(r1)
      : "r"(lo), "r"(hi), "r"(a1));
  (void)t0;
  return (uint64_t)r0 | ((uint64_t)r1 << 32);
}

// The same recombination for accumulators of up to 24 bits (K = 256 bytes per row: 256 * 255^2 < 2^24, the
// reconstruction kernels with more than 16 Fp61 shares per row).  There p45 * 2^32 alone can reach 2^64, so its bits
// from 29 up are folded first (2^61 = 1): R < 2^32 + 2^48 + 2^61 + 2^61 + 2^19 < 2^63 for ANY limb values below 2^24.
__device__ __forceinline__ uint64_t tc_combine24(const uint32_t* v) {
  const uint32_t p01 = v[0] + (v[1] << 8), p23 = v[2] + (v[3] << 8);  // < 2^32 (v < 2^24 - 2^17)
  const uint32_t p45 = v[4] + (v[5] << 8), p67 = v[6] + (v[7] << 8);
  const uint64_t R = (uint64_t)p01 + ((uint64_t)p23 << 16) + ((uint64_t)(p45 & 0x1FFFFFFFu) << 32) + (uint64_t)(p45 >> 29) +
                     ((uint64_t)(p67 & 0x1FFFu) << 48) + (uint64_t)(p67 >> 13);
  return F61::from_raw(R);
}

__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Fp127 (mersenne127.cc:60-97 semantics): sixteen 23-bit limbs at 2^(8s) -> canonical residue mod 2^127 - 1.
// Word-level: pairs q_m = v_2m + v_2m+1 * 2^8 (< 2^32, weight 2^(16m)); the even pairs concatenate into a
// 128-bit number E, the odd ones into O with weight 2^16; X = E + (O << 16) is formed in five 32-bit words
// with one carry chain, folded once at bit 127 (2^127 = 1), once more for the single possible carry, and
// p itself is mapped to 0.
__device__ __forceinline__ E127 tc_combine127(const uint32_t* v) {
  uint32_t q[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) q[m] = v[2 * m] + (v[2 * m + 1] << 8);
  // W = O << 16 as words w0..w4 (O = q1 | q3<<32 | q5<<64 | q7<<96)
  const uint32_t w0 = q[1] << 16;
  const uint32_t w1 = __funnelshift_l(q[1], q[3], 16);
  const uint32_t w2 = __funnelshift_l(q[3], q[5], 16);
  const uint32_t w3 = __funnelshift_l(q[5], q[7], 16);
  const uint32_t w4 = q[7] >> 16;
  uint32_t s0, s1, s2, s3, s4;  // X = E + W, E = q0 | q2<<32 | q4<<64 | q6<<96
  asm("add.cc.u32 %0, %5, %9;\n\t"
      "addc.cc.u32 %1, %6, %10;\n\t"
      "addc.cc.u32 %2, %7, %11;\n\t"
      "addc.cc.u32 %3, %8, %12;\n\t"
      "addc.u32 %4, %13, 0;"
      : "=r"(s0), "=r"(s1), "=r"(s2), "=r"(s3), "=r"(s4)
      : "r"(q[0]), "r"(q[2]), "r"(q[4]), "r"(q[6]), "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(w4));
  // fold bits >= 127: hi = X >> 127 (< 2^18)
  const uint32_t hi = __funnelshift_l(s3, s4, 1);
  s3 &= 0x7FFFFFFFu;
  asm("add.cc.u32 %0, %0, %4;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.u32 %3, %3, 0;"
      : "+r"(s0), "+r"(s1), "+r"(s2), "+r"(s3)
      : "r"(hi));
  // now < 2^127 + 2^18: if bit 127 is set the rest is < 2^18, so adding the carry cannot ripple
  s0 += s3 >> 31;
  s3 &= 0x7FFFFFFFu;
  const bool is_p = (s0 & s1 & s2 & (s3 | 0x80000000u)) == 0xFFFFFFFFu;
  E127 r;
  r.lo = is_p ? 0 : ((uint64_t)s0 | ((uint64_t)s1 << 32));
  r.hi = is_p ? 0 : ((uint64_t)s2 | ((uint64_t)s3 << 32));
  return r;
}

}  // namespace sclgpu
