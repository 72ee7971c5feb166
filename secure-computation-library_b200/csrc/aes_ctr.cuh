// aes_ctr.cuh -- AES-128-CTR keystream of scl::util::PRG on sm_100a.
//
// Reference behaviour (src/scl/util/prg.cc:82-84, 124-146; include/scl/util/prg.h:34-43):
//   block i = AES128_seed( LE64(i) || LE64(PRG_NONCE) ),  PRG_NONCE = 0x0123456789ABCDEF,
//   counter starts at PRG_INITIAL_COUNTER = 0 and advances by one per block.
// The reference uses AES-NI; here AES-128 (FIPS-197) is done with four T-tables
// in shared memory.  Layout (DESIGN.md "AES tables"): every table is replicated
// once per lane so that lane l only ever touches bank l -- conflict-free for any
// index pattern -- and the row stride is 256 bytes so that ONE byte-permute
// (PRMT) builds a lookup address from a state byte:
//
//   byte address = tbase + idx*256 + (tbl&1)*128 + lane*4 + (tbl>>1)*65536
//                  ^^^^^ 64 KiB aligned: address byte 1 is exactly idx
//
// State/round-key words are little-endian column words (byte 0 = row 0).
#pragma once
#include <cstdint>

namespace sclgpu {

struct AesKey {
  uint32_t rk[44];  // 11 round keys x 4 LE column words (host-expanded, prg.cc:54-75)
  // 2^8, 2^16, 2^24 as RUN-TIME values (set by the host): multiplying by them moves byte
  // extraction from the ALU pipe (PRMT) to the FMA pipe (IMAD / IMAD.HI), see aes_addr_fma
  uint32_t k8, k16, k24;
};

// Measured on B200 (2^26 secrets, n=32, t=15): share 11.1 ms with PRMT addresses, 12.9 ms with the
// FMA-pipe form below (IMAD.HI is a multi-issue instruction); keystream 2.11 vs 2.40 ms.  Off.
#ifndef SCLGPU_AES_FMA_ADDR
#define SCLGPU_AES_FMA_ADDR 0
#endif
// final round through 8-bit loads of S[x] and FMA-pipe joins (1) or T-table words joined by byte-permutes (0)
#ifndef SCLGPU_AES_LAST_U8
#define SCLGPU_AES_LAST_U8 1
#endif

static constexpr uint32_t kPrgNonceLo = 0x89ABCDEFu;  // PRG_NONCE, prg.h:34-36
static constexpr uint32_t kPrgNonceHi = 0x01234567u;
static constexpr uint32_t kAesTableBytes = 128u * 1024u;  // 4 tables x 256 rows x 32 lanes x 4 B
static constexpr uint32_t kAesTableAlign = 64u * 1024u;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

// First 64 KiB-aligned shared address at or after `dyn` (the dynamic smem base).
__device__ __forceinline__ uint32_t aes_table_base(const void* dyn) {
  return (smem_u32(dyn) + (kAesTableAlign - 1)) & ~(kAesTableAlign - 1);
}

// Cooperative fill of the replicated tables from T0 (256 words in global memory):
// T0[x] = (2s, s, s, 3s) LE, T1/T2/T3 = T0 rotated left by 8/16/24 bits.
__device__ __forceinline__ void aes_fill_tables(uint32_t tbase, const uint32_t* __restrict__ g_t0) {
  for (uint32_t e = threadIdx.x; e < 256u * 32u; e += blockDim.x) {
    const uint32_t idx = e >> 5, l = e & 31u;
    const uint32_t t0 = __ldg(g_t0 + idx);
    const uint32_t t1 = __funnelshift_l(t0, t0, 8);
    const uint32_t t2 = __funnelshift_l(t0, t0, 16);
    const uint32_t t3 = __funnelshift_l(t0, t0, 24);
    const uint32_t a = tbase + idx * 256u + l * 4u;
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(t0) : "memory");
    asm volatile("st.shared.u32 [%0+128], %1;" ::"r"(a), "r"(t1) : "memory");
    asm volatile("st.shared.u32 [%0+65536], %1;" ::"r"(a), "r"(t2) : "memory");
    asm volatile("st.shared.u32 [%0+65664], %1;" ::"r"(a), "r"(t3) : "memory");
  }
}

template <int OFF>
__device__ __forceinline__ uint32_t aes_lds(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(addr), "n"(OFF));
  return v;
}

// lookup address for byte K of w: (lanebase with byte 1 replaced by that byte)
template <int K>
__device__ __forceinline__ uint32_t aes_addr(uint32_t w, uint32_t lanebase) {
  return __byte_perm(w, lanebase, 0x7604 | (K << 4));
}

// One AES-128 block.  lanebase = tbase + lane*4 (byte 1 zero).
__device__ __forceinline__ void aes128_encrypt(const AesKey& key, uint32_t lanebase, uint32_t s0,
                                               uint32_t s1, uint32_t s2, uint32_t s3,
                                               uint32_t& o0, uint32_t& o1, uint32_t& o2,
                                               uint32_t& o3) {
  s0 ^= key.rk[0];
  s1 ^= key.rk[1];
  s2 ^= key.rk[2];
  s3 ^= key.rk[3];
#pragma unroll
  for (int r = 1; r < 10; ++r) {
    const uint32_t a00 = aes_addr<0>(s0, lanebase), a01 = aes_addr<1>(s0, lanebase),
                   a02 = aes_addr<2>(s0, lanebase), a03 = aes_addr<3>(s0, lanebase);
    const uint32_t a10 = aes_addr<0>(s1, lanebase), a11 = aes_addr<1>(s1, lanebase),
                   a12 = aes_addr<2>(s1, lanebase), a13 = aes_addr<3>(s1, lanebase);
    const uint32_t a20 = aes_addr<0>(s2, lanebase), a21 = aes_addr<1>(s2, lanebase),
                   a22 = aes_addr<2>(s2, lanebase), a23 = aes_addr<3>(s2, lanebase);
    const uint32_t a30 = aes_addr<0>(s3, lanebase), a31 = aes_addr<1>(s3, lanebase),
                   a32 = aes_addr<2>(s3, lanebase), a33 = aes_addr<3>(s3, lanebase);
    // column j = T0[b0(s_j)] ^ T1[b1(s_j+1)] ^ T2[b2(s_j+2)] ^ T3[b3(s_j+3)] ^ rk
    const uint32_t t0 = aes_lds<0>(a00) ^ aes_lds<128>(a11) ^ aes_lds<65536>(a22) ^
                        aes_lds<65664>(a33) ^ key.rk[4 * r + 0];
    const uint32_t t1 = aes_lds<0>(a10) ^ aes_lds<128>(a21) ^ aes_lds<65536>(a32) ^
                        aes_lds<65664>(a03) ^ key.rk[4 * r + 1];
    const uint32_t t2 = aes_lds<0>(a20) ^ aes_lds<128>(a31) ^ aes_lds<65536>(a02) ^
                        aes_lds<65664>(a13) ^ key.rk[4 * r + 2];
    const uint32_t t3 = aes_lds<0>(a30) ^ aes_lds<128>(a01) ^ aes_lds<65536>(a12) ^
                        aes_lds<65664>(a23) ^ key.rk[4 * r + 3];
    s0 = t0;
    s1 = t1;
    s2 = t2;
    s3 = t3;
  }
  // final round: SubBytes + ShiftRows + AddRoundKey.  S[x] sits in byte 0 of T2/T3,
  // byte 1 of T0/T3, byte 2 of T0/T1, byte 3 of T1/T2.
  {
    const uint32_t a00 = aes_addr<0>(s0, lanebase), a01 = aes_addr<1>(s0, lanebase),
                   a02 = aes_addr<2>(s0, lanebase), a03 = aes_addr<3>(s0, lanebase);
    const uint32_t a10 = aes_addr<0>(s1, lanebase), a11 = aes_addr<1>(s1, lanebase),
                   a12 = aes_addr<2>(s1, lanebase), a13 = aes_addr<3>(s1, lanebase);
    const uint32_t a20 = aes_addr<0>(s2, lanebase), a21 = aes_addr<1>(s2, lanebase),
                   a22 = aes_addr<2>(s2, lanebase), a23 = aes_addr<3>(s2, lanebase);
    const uint32_t a30 = aes_addr<0>(s3, lanebase), a31 = aes_addr<1>(s3, lanebase),
                   a32 = aes_addr<2>(s3, lanebase), a33 = aes_addr<3>(s3, lanebase);
#define SCLGPU_AES_LAST(x0, x1, x2, x3)                                               \
  __byte_perm(__byte_perm(aes_lds<65536>(x0), aes_lds<65664>(x1), 0x0050),            \
              __byte_perm(aes_lds<0>(x2), aes_lds<128>(x3), 0x7200), 0x7610)
    o0 = SCLGPU_AES_LAST(a00, a11, a22, a33) ^ key.rk[40];
    o1 = SCLGPU_AES_LAST(a10, a21, a32, a03) ^ key.rk[41];
    o2 = SCLGPU_AES_LAST(a20, a31, a02, a13) ^ key.rk[42];
    o3 = SCLGPU_AES_LAST(a30, a01, a12, a23) ^ key.rk[43];
#undef SCLGPU_AES_LAST
  }
}

// ---------------------------------------------------------------------------
// Counter-mode caching.  The plaintext of block `ctr` is LE64(ctr) || LE64(nonce):
// only state byte 0 (the low counter byte) differs between blocks that share
// ctr >> 8.  Byte 0 feeds one column of round 1 (t0), and t0 feeds one T-table
// term of each column of round 2, so everything else of rounds 1-2 is computed
// once per 256-block group (27 lookups) and each block costs 1 + 4 lookups for
// rounds 1-2 instead of 32.  Values are bit-identical to aes128_encrypt.
struct PrgGroup {
  uint32_t k0;              // round-1 column 0 without its T0 term
  uint32_t u0, u1, u2, u3;  // round-2 columns without their t0 terms
};

// T-table lookup of byte K of w in table TBL
template <int TBL, int K>
__device__ __forceinline__ uint32_t aes_t(uint32_t w, uint32_t lanebase) {
  constexpr int OFF = (TBL & 1) * 128 + (TBL >> 1) * 65536;
  return aes_lds<OFF>(aes_addr<K>(w, lanebase));
}

// The same lookup address, formed on the FMA pipe for the two end bytes (the ALU pipe is the
// co-critical pipe of the fused kernels, the FMA pipe is idle):
//   byte 3:  lanebase + (w >> 24) * 256  =  mad.lo(mul.hi(w, 2^8), 2^8, lanebase)
//   byte 0:  lanebase + (w & 255) * 256  =  mad.hi(w * 2^24, 2^16, lanebase)
// bytes 1 and 2 would need three multiplies each and stay on PRMT.
template <int K>
__device__ __forceinline__ uint32_t aes_addr_fma(const AesKey& key, uint32_t w, uint32_t lanebase) {
#if SCLGPU_AES_FMA_ADDR
  if (K == 3) {
    uint32_t t, a;
    asm("mul.hi.u32 %0, %1, %2;" : "=r"(t) : "r"(w), "r"(key.k8));
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(t), "r"(key.k8), "r"(lanebase));
    return a;
  }
  if (K == 0) {
    uint32_t t, a;
    asm("mul.lo.u32 %0, %1, %2;" : "=r"(t) : "r"(w), "r"(key.k24));
    asm("mad.hi.u32 %0, %1, %2, %3;" : "=r"(a) : "r"(t), "r"(key.k16), "r"(lanebase));
    return a;
  }
#endif
  return aes_addr<K>(w, lanebase);
}
template <int TBL, int K>
__device__ __forceinline__ uint32_t aes_tk(const AesKey& key, uint32_t w, uint32_t lanebase) {
  constexpr int OFF = (TBL & 1) * 128 + (TBL >> 1) * 65536;
  return aes_lds<OFF>(aes_addr_fma<K>(key, w, lanebase));
}

// group state for the 256 counters sharing ctr >> 8
__device__ __forceinline__ void prg_group(const AesKey& key, uint32_t lanebase, uint64_t ctr,
                                          PrgGroup& g) {
  const uint32_t a0 = (uint32_t)ctr ^ key.rk[0], a1 = (uint32_t)(ctr >> 32) ^ key.rk[1],
                 a2 = kPrgNonceLo ^ key.rk[2], a3 = kPrgNonceHi ^ key.rk[3];
  g.k0 = aes_t<1, 1>(a1, lanebase) ^ aes_t<2, 2>(a2, lanebase) ^ aes_t<3, 3>(a3, lanebase) ^ key.rk[4];
  const uint32_t t1 = aes_t<0, 0>(a1, lanebase) ^ aes_t<1, 1>(a2, lanebase) ^ aes_t<2, 2>(a3, lanebase) ^
                      aes_t<3, 3>(a0, lanebase) ^ key.rk[5];
  const uint32_t t2 = aes_t<0, 0>(a2, lanebase) ^ aes_t<1, 1>(a3, lanebase) ^ aes_t<2, 2>(a0, lanebase) ^
                      aes_t<3, 3>(a1, lanebase) ^ key.rk[6];
  const uint32_t t3 = aes_t<0, 0>(a3, lanebase) ^ aes_t<1, 1>(a0, lanebase) ^ aes_t<2, 2>(a1, lanebase) ^
                      aes_t<3, 3>(a2, lanebase) ^ key.rk[7];
  g.u0 = aes_t<1, 1>(t1, lanebase) ^ aes_t<2, 2>(t2, lanebase) ^ aes_t<3, 3>(t3, lanebase) ^ key.rk[8];
  g.u1 = aes_t<0, 0>(t1, lanebase) ^ aes_t<1, 1>(t2, lanebase) ^ aes_t<2, 2>(t3, lanebase) ^ key.rk[9];
  g.u2 = aes_t<0, 0>(t2, lanebase) ^ aes_t<1, 1>(t3, lanebase) ^ aes_t<3, 3>(t1, lanebase) ^ key.rk[10];
  g.u3 = aes_t<0, 0>(t3, lanebase) ^ aes_t<2, 2>(t1, lanebase) ^ aes_t<3, 3>(t2, lanebase) ^ key.rk[11];
}

// The group state again, for a thread whose consecutive groups share the counter bits from 16 up (consecutive tiles of a
// contiguous range of secrets): of the 27 lookups only the ones fed by counter byte 1 change from group to group --
// the T1 term of round-1 column 3 and the four round-2 terms that column feeds.  Everything else is kept per thread and
// rebuilt when ctr >> 16 changes (every 65536 blocks): 5 lookups per group instead of 27.  Same values as prg_group.
struct PrgGroupCache {
  uint64_t tag;                 // ctr >> 16 the cached words belong to (~0: none)
  uint32_t k0, q3, p0, p1, p2, p3;
};

__device__ __forceinline__ void prg_group_cached(const AesKey& key, uint32_t lanebase, uint64_t ctr, PrgGroup& g,
                                                 PrgGroupCache& c) {
  const uint32_t a0 = (uint32_t)ctr ^ key.rk[0];
  if ((ctr >> 16) != c.tag) {
    c.tag = ctr >> 16;
    const uint32_t a1 = (uint32_t)(ctr >> 32) ^ key.rk[1], a2 = kPrgNonceLo ^ key.rk[2], a3 = kPrgNonceHi ^ key.rk[3];
    c.k0 = aes_t<1, 1>(a1, lanebase) ^ aes_t<2, 2>(a2, lanebase) ^ aes_t<3, 3>(a3, lanebase) ^ key.rk[4];
    const uint32_t t1 = aes_t<0, 0>(a1, lanebase) ^ aes_t<1, 1>(a2, lanebase) ^ aes_t<2, 2>(a3, lanebase) ^
                        aes_t<3, 3>(a0, lanebase) ^ key.rk[5];
    const uint32_t t2 = aes_t<0, 0>(a2, lanebase) ^ aes_t<1, 1>(a3, lanebase) ^ aes_t<2, 2>(a0, lanebase) ^
                        aes_t<3, 3>(a1, lanebase) ^ key.rk[6];
    c.q3 = aes_t<0, 0>(a3, lanebase) ^ aes_t<2, 2>(a1, lanebase) ^ aes_t<3, 3>(a2, lanebase) ^ key.rk[7];
    c.p0 = aes_t<1, 1>(t1, lanebase) ^ aes_t<2, 2>(t2, lanebase) ^ key.rk[8];
    c.p1 = aes_t<0, 0>(t1, lanebase) ^ aes_t<1, 1>(t2, lanebase) ^ key.rk[9];
    c.p2 = aes_t<0, 0>(t2, lanebase) ^ aes_t<3, 3>(t1, lanebase) ^ key.rk[10];
    c.p3 = aes_t<2, 2>(t1, lanebase) ^ aes_t<3, 3>(t2, lanebase) ^ key.rk[11];
  }
  const uint32_t t3 = c.q3 ^ aes_t<1, 1>(a0, lanebase);
  g.k0 = c.k0;
  g.u0 = c.p0 ^ aes_t<3, 3>(t3, lanebase);
  g.u1 = c.p1 ^ aes_t<2, 2>(t3, lanebase);
  g.u2 = c.p2 ^ aes_t<1, 1>(t3, lanebase);
  g.u3 = c.p3 ^ aes_t<0, 0>(t3, lanebase);
}

// S[byte K of w], zero-extended: byte 1 of the lane's T0 entry (T0[x] = (2s, s, s, 3s))
template <int K>
__device__ __forceinline__ uint32_t aes_sbox(uint32_t w, uint32_t lanebase) {
  uint32_t v;
  asm("ld.shared.u8 %0, [%1+1];" : "=r"(v) : "r"(aes_addr<K>(w, lanebase)));
  return v;
}
// b0 | b1 << 8 | b2 << 16 | b3 << 24 for bytes b* < 256, as three IMADs (k8, k16: 2^8, 2^16 as run-time values)
__device__ __forceinline__ uint32_t aes_join4(const AesKey& key, uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
  uint32_t lo, hi, r;
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(lo) : "r"(b1), "r"(key.k8), "r"(b0));
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(hi) : "r"(b3), "r"(key.k8), "r"(b2));
  asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(hi), "r"(key.k16), "r"(lo));
  return r;
}

// rounds R0..9 and the final round on state (s0..s3) = output of round R0-1
// low == false: output words 0 and 1 are not wanted (keystream block 0 of an Fp61 sharing: slot 0 of the draw is
// replaced by the secret, shamir.h:56-57) -- their eight final-round lookups are skipped and o0, o1 are left untouched
template <int R0>
__device__ __forceinline__ void aes128_tail(const AesKey& key, uint32_t lanebase, uint32_t s0, uint32_t s1,
                                            uint32_t s2, uint32_t s3, uint32_t& o0, uint32_t& o1,
                                            uint32_t& o2, uint32_t& o3, bool low = true) {
#pragma unroll
  for (int r = R0; r < 10; ++r) {
    const uint32_t t0 = aes_tk<0, 0>(key, s0, lanebase) ^ aes_tk<1, 1>(key, s1, lanebase) ^ aes_tk<2, 2>(key, s2, lanebase) ^
                        aes_tk<3, 3>(key, s3, lanebase) ^ key.rk[4 * r + 0];
    const uint32_t t1 = aes_tk<0, 0>(key, s1, lanebase) ^ aes_tk<1, 1>(key, s2, lanebase) ^ aes_tk<2, 2>(key, s3, lanebase) ^
                        aes_tk<3, 3>(key, s0, lanebase) ^ key.rk[4 * r + 1];
    const uint32_t t2 = aes_tk<0, 0>(key, s2, lanebase) ^ aes_tk<1, 1>(key, s3, lanebase) ^ aes_tk<2, 2>(key, s0, lanebase) ^
                        aes_tk<3, 3>(key, s1, lanebase) ^ key.rk[4 * r + 2];
    const uint32_t t3 = aes_tk<0, 0>(key, s3, lanebase) ^ aes_tk<1, 1>(key, s0, lanebase) ^ aes_tk<2, 2>(key, s1, lanebase) ^
                        aes_tk<3, 3>(key, s2, lanebase) ^ key.rk[4 * r + 3];
    s0 = t0;
    s1 = t1;
    s2 = t2;
    s3 = t3;
  }
#if SCLGPU_AES_LAST_U8
  // final round (SubBytes + ShiftRows + AddRoundKey): S[x] is byte 1 of the T0 entry, read zero-extended by an 8-bit
  // load; the four bytes of an output word are joined by integer multiply-adds with RUN-TIME multipliers (FMA pipe),
  // so the ALU pipe -- co-critical with the shared-memory pipe in the fused kernels -- sees one XOR per word
  // instead of three byte-permutes and the XOR.
#define SCLGPU_AES_LAST2(w0, w1, w2, w3)                                                                   \
  aes_join4(key, aes_sbox<0>(w0, lanebase), aes_sbox<1>(w1, lanebase), aes_sbox<2>(w2, lanebase), aes_sbox<3>(w3, lanebase))
#else
  // final round: S[x] sits in byte 0 of T2/T3, byte 1 of T0/T3, byte 2 of T0/T1, byte 3 of T1/T2
#define SCLGPU_AES_LAST2(w0, w1, w2, w3)                                                          \
  __byte_perm(__byte_perm(aes_tk<2, 0>(key, w0, lanebase), aes_tk<3, 1>(key, w1, lanebase), 0x0050),          \
              __byte_perm(aes_tk<0, 2>(key, w2, lanebase), aes_tk<1, 3>(key, w3, lanebase), 0x7200), 0x7610)
#endif
  if (low) {
    o0 = SCLGPU_AES_LAST2(s0, s1, s2, s3) ^ key.rk[40];
    o1 = SCLGPU_AES_LAST2(s1, s2, s3, s0) ^ key.rk[41];
  }
  o2 = SCLGPU_AES_LAST2(s2, s3, s0, s1) ^ key.rk[42];
  o3 = SCLGPU_AES_LAST2(s3, s0, s1, s2) ^ key.rk[43];
#undef SCLGPU_AES_LAST2
}

// keystream block whose counter has low 32 bits ctr_lo and lies in group g
__device__ __forceinline__ void prg_block_grouped(const AesKey& key, uint32_t lanebase, const PrgGroup& g,
                                                  uint32_t ctr_lo, uint32_t& o0, uint32_t& o1, uint32_t& o2,
                                                  uint32_t& o3, bool low = true) {
  const uint32_t t0 = aes_t<0, 0>(ctr_lo ^ key.rk[0], lanebase) ^ g.k0;
  const uint32_t s0 = aes_tk<0, 0>(key, t0, lanebase) ^ g.u0;
  const uint32_t s1 = aes_tk<3, 3>(key, t0, lanebase) ^ g.u1;
  const uint32_t s2 = aes_t<2, 2>(t0, lanebase) ^ g.u2;
  const uint32_t s3 = aes_t<1, 1>(t0, lanebase) ^ g.u3;
  aes128_tail<3>(key, lanebase, s0, s1, s2, s3, o0, o1, o2, o3, low);
}

// keystream block `ctr` of the PRG (prg.cc:82-84): plaintext = LE64(ctr) || LE64(nonce)
__device__ __forceinline__ void prg_block(const AesKey& key, uint32_t lanebase, uint64_t ctr,
                                          uint32_t& o0, uint32_t& o1, uint32_t& o2, uint32_t& o3) {
  aes128_encrypt(key, lanebase, (uint32_t)ctr, (uint32_t)(ctr >> 32), kPrgNonceLo, kPrgNonceHi, o0,
                 o1, o2, o3);
}

}  // namespace sclgpu
