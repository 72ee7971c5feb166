// aes_bitsliced.cuh -- the AES-128-CTR keystream of scl::util::PRG (src/scl/util/prg.cc:82-84, 124-146) WITHOUT table
// lookups: bitsliced, 32 blocks per thread.  This is the measured comparison arm for the T-table kernels of aes_ctr.cuh
// (north_star: "a bitsliced or T-table AES-CTR keystream"): selected with SCLGPU_PRG_BITSLICED=1 for
// sclgpu_prg_expand[_dev], bit-identical output, checked against the oracle on the CPU (tests/cpp/bitsliced_check.cc)
// and on the GPU.
//
// Representation: s[8 * i + k] holds bit k (0 = least significant) of state byte i (FIPS-197 order: byte i = row i % 4
// of column i / 4 = byte i of the 16-byte block) for 32 blocks at once -- bit l of every word belongs to block
// ctr0 + l.  Counter-mode input needs no transposition: with ctr0 a multiple of 32 the five low counter bits are the
// constants 0xAAAAAAAA, 0xCCCCCCCC, ..., every other plaintext bit is the same in all 32 blocks.
//   SubBytes    : the 113-gate circuit of Boyar and Peralta ("A new combinational logic minimization technique with
//                 applications to cryptology", 2010: 32 AND, 77 XOR, 4 XNOR), sixteen times per round on eight words
//   ShiftRows   : register moves (the round body is not unrolled over rounds, so positions stay fixed)
//   MixColumns  : out_r = xtime(a_r ^ a_{r+1}) ^ a_r ^ (a_0 ^ a_1 ^ a_2 ^ a_3) on planes, in place per column
//   AddRoundKey : XOR with 0 / ~0 masks, one per key bit (11 x 128 words, built once per CTA in shared memory)
// and a 32 x 32 bit transposition per output word at the end.  About 2 250 logic instructions per round and thread
// = 70 per block and round, against 16 lookups + 21 ALU instructions of the T-table form.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SCLGPU_BS_HD __host__ __device__ __forceinline__
#else
#define SCLGPU_BS_HD inline
#endif

namespace sclgpu {

static constexpr uint32_t kBsKeyWords = 11u * 128u;  // one mask per bit of the eleven round keys

// round-key masks from the expanded key (44 little-endian column words, prg.cc:54-75): mask (r, i, k) = all ones iff
// bit k of byte i of round key r is set
SCLGPU_BS_HD uint32_t bs_key_mask(const uint32_t* rk, uint32_t idx) {
  const uint32_t r = idx >> 7, p = idx & 127u, i = p >> 3, k = p & 7u;
  return ((rk[4u * r + (i >> 2)] >> (8u * (i & 3u) + k)) & 1u) ? 0xFFFFFFFFu : 0u;
}

// S-box on the eight planes of one byte, in place.  u[7] is the most significant bit (the circuit's U0).
SCLGPU_BS_HD void bs_sbox(uint32_t* u) {
  const uint32_t U0 = u[7], U1 = u[6], U2 = u[5], U3 = u[4], U4 = u[3], U5 = u[2], U6 = u[1], U7 = u[0];
  const uint32_t T1 = U0 ^ U3, T2 = U0 ^ U5, T3 = U0 ^ U6, T4 = U3 ^ U5, T5 = U4 ^ U6, T6 = T1 ^ T5, T7 = U1 ^ U2;
  const uint32_t T8 = U7 ^ T6, T9 = U7 ^ T7, T10 = T6 ^ T7, T11 = U1 ^ U5, T12 = U2 ^ U5, T13 = T3 ^ T4, T14 = T6 ^ T11;
  const uint32_t T15 = T5 ^ T11, T16 = T5 ^ T12, T17 = T9 ^ T16, T18 = U3 ^ U7, T19 = T7 ^ T18, T20 = T1 ^ T19;
  const uint32_t T21 = U6 ^ U7, T22 = T7 ^ T21, T23 = T2 ^ T22, T24 = T2 ^ T10, T25 = T20 ^ T17, T26 = T3 ^ T16;
  const uint32_t T27 = T1 ^ T12;
  const uint32_t M1 = T13 & T6, M2 = T23 & T8, M3 = T14 ^ M1, M4 = T19 & U7, M5 = M4 ^ M1, M6 = T3 & T16, M7 = T22 & T9;
  const uint32_t M8 = T26 ^ M6, M9 = T20 & T17, M10 = M9 ^ M6, M11 = T1 & T15, M12 = T4 & T27, M13 = M12 ^ M11;
  const uint32_t M14 = T2 & T10, M15 = M14 ^ M11, M16 = M3 ^ M2, M17 = M5 ^ T24, M18 = M8 ^ M7, M19 = M10 ^ M15;
  const uint32_t M20 = M16 ^ M13, M21 = M17 ^ M15, M22 = M18 ^ M13, M23 = M19 ^ T25, M24 = M22 ^ M23, M25 = M22 & M20;
  const uint32_t M26 = M21 ^ M25, M27 = M20 ^ M21, M28 = M23 ^ M25, M29 = M28 & M27, M30 = M26 & M24, M31 = M20 & M23;
  const uint32_t M32 = M27 & M31, M33 = M27 ^ M25, M34 = M21 & M22, M35 = M24 & M34, M36 = M24 ^ M25, M37 = M21 ^ M29;
  const uint32_t M38 = M32 ^ M33, M39 = M23 ^ M30, M40 = M35 ^ M36, M41 = M38 ^ M40, M42 = M37 ^ M39, M43 = M37 ^ M38;
  const uint32_t M44 = M39 ^ M40, M45 = M42 ^ M41;
  const uint32_t M46 = M44 & T6, M47 = M40 & T8, M48 = M39 & U7, M49 = M43 & T16, M50 = M38 & T9, M51 = M37 & T17;
  const uint32_t M52 = M42 & T15, M53 = M45 & T27, M54 = M41 & T10, M55 = M44 & T13, M56 = M40 & T23, M57 = M39 & T19;
  const uint32_t M58 = M43 & T3, M59 = M38 & T22, M60 = M37 & T20, M61 = M42 & T1, M62 = M45 & T4, M63 = M41 & T2;
  const uint32_t L0 = M61 ^ M62, L1 = M50 ^ M56, L2 = M46 ^ M48, L3 = M47 ^ M55, L4 = M54 ^ M58, L5 = M49 ^ M61;
  const uint32_t L6 = M62 ^ L5, L7 = M46 ^ L3, L8 = M51 ^ M59, L9 = M52 ^ M53, L10 = M53 ^ L4, L11 = M60 ^ L2;
  const uint32_t L12 = M48 ^ M51, L13 = M50 ^ L0, L14 = M52 ^ M61, L15 = M55 ^ L1, L16 = M56 ^ L0, L17 = M57 ^ L1;
  const uint32_t L18 = M58 ^ L8, L19 = M63 ^ L4, L20 = L0 ^ L1, L21 = L1 ^ L7, L22 = L3 ^ L12, L23 = L18 ^ L2;
  const uint32_t L24 = L15 ^ L9, L25 = L6 ^ L10, L26 = L7 ^ L9, L27 = L8 ^ L10, L28 = L11 ^ L14, L29 = L11 ^ L17;
  u[7] = L6 ^ L24;
  u[6] = ~(L16 ^ L26);
  u[5] = ~(L19 ^ L28);
  u[4] = L6 ^ L21;
  u[3] = L20 ^ L22;
  u[2] = L25 ^ L29;
  u[1] = ~(L13 ^ L27);
  u[0] = ~(L6 ^ L23);
}

// ShiftRows: row r of the state (bytes r, r + 4, r + 8, r + 12) rotates left by r columns
SCLGPU_BS_HD void bs_shift_rows(uint32_t (&s)[128]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    uint32_t t = s[8 * 1 + k];  // row 1: by one
    s[8 * 1 + k] = s[8 * 5 + k];
    s[8 * 5 + k] = s[8 * 9 + k];
    s[8 * 9 + k] = s[8 * 13 + k];
    s[8 * 13 + k] = t;
    t = s[8 * 2 + k];  // row 2: by two
    s[8 * 2 + k] = s[8 * 10 + k];
    s[8 * 10 + k] = t;
    t = s[8 * 6 + k];
    s[8 * 6 + k] = s[8 * 14 + k];
    s[8 * 14 + k] = t;
    t = s[8 * 15 + k];  // row 3: by three = right by one
    s[8 * 15 + k] = s[8 * 11 + k];
    s[8 * 11 + k] = s[8 * 7 + k];
    s[8 * 7 + k] = s[8 * 3 + k];
    s[8 * 3 + k] = t;
  }
}

// MixColumns on column c (bytes 4c .. 4c + 3), in place
SCLGPU_BS_HD void bs_mix_column(uint32_t* a) {
  uint32_t t[4][8], sum[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    t[0][k] = a[k] ^ a[8 + k];
    t[1][k] = a[8 + k] ^ a[16 + k];
    t[2][k] = a[16 + k] ^ a[24 + k];
    t[3][k] = a[24 + k] ^ a[k];
    sum[k] = t[0][k] ^ t[2][k];
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    // xtime on planes: (b7, b0 ^ b7, b1, b2 ^ b7, b3 ^ b7, b4, b5, b6), least significant first
    const uint32_t* x = t[r];
    uint32_t* o = a + 8 * r;
    const uint32_t h = x[7];
    o[0] ^= sum[0] ^ h;
    o[1] ^= sum[1] ^ x[0] ^ h;
    o[2] ^= sum[2] ^ x[1];
    o[3] ^= sum[3] ^ x[2] ^ h;
    o[4] ^= sum[4] ^ x[3] ^ h;
    o[5] ^= sum[5] ^ x[4];
    o[6] ^= sum[6] ^ x[5];
    o[7] ^= sum[7] ^ x[6];
  }
}

// planes of the 32 plaintext blocks LE64(ctr0 + l) || LE64(PRG_NONCE), l = 0..31 (prg.cc:82-84); ctr0 % 32 == 0
SCLGPU_BS_HD void bs_ctr_planes(uint64_t ctr0, uint32_t (&s)[128]) {
  const uint64_t nonce = 0x0123456789ABCDEFull;  // PRG_NONCE, prg.h:34-36
#pragma unroll
  for (int p = 0; p < 64; ++p) {
    s[p] = ((ctr0 >> p) & 1u) ? 0xFFFFFFFFu : 0u;
    s[64 + p] = ((nonce >> p) & 1u) ? 0xFFFFFFFFu : 0u;
  }
  s[0] = 0xAAAAAAAAu;
  s[1] = 0xCCCCCCCCu;
  s[2] = 0xF0F0F0F0u;
  s[3] = 0xFF00FF00u;
  s[4] = 0xFFFF0000u;
}

// in-place transposition of a 32 x 32 bit matrix: afterwards bit p of a[l] = bit l of the old a[p]
SCLGPU_BS_HD void bs_transpose32(uint32_t* a) {
  uint32_t m = 0x0000FFFFu;
#pragma unroll
  for (int j = 16; j != 0; j >>= 1, m ^= (m << j)) {
#pragma unroll
    for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
      const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
      a[k] ^= t << j;
      a[k + j] ^= t;
    }
  }
}

// 32 keystream blocks: on return s[32 * w + l] is little-endian word w of block ctr0 + l.  km: kBsKeyWords masks.
SCLGPU_BS_HD void bs_aes_ctr32(const uint32_t* km, uint64_t ctr0, uint32_t (&s)[128]) {
  bs_ctr_planes(ctr0, s);
#pragma unroll
  for (int p = 0; p < 128; ++p) s[p] ^= km[p];
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
  for (int r = 1; r <= 10; ++r) {
#pragma unroll
    for (int i = 0; i < 16; ++i) bs_sbox(s + 8 * i);
    bs_shift_rows(s);
    if (r < 10) {
#pragma unroll
      for (int c = 0; c < 4; ++c) bs_mix_column(s + 32 * c);
    }
    const uint32_t* kr = km + 128 * r;
#pragma unroll
    for (int p = 0; p < 128; ++p) s[p] ^= kr[p];
  }
#pragma unroll
  for (int w = 0; w < 4; ++w) bs_transpose32(s + 32 * w);
}

#if defined(__CUDACC__)
// PRG::next keystream, blocks [first_block, first_block + n_blocks) into out (16-byte aligned, whole blocks).  A thread
// takes one aligned group of 32 counters at a time and stores the blocks of it that are inside the range.
__global__ void __launch_bounds__(128, 1)
k_prg_bitsliced(const __grid_constant__ AesKey key, uint64_t first_block, uint64_t n_blocks, uint4* __restrict__ out) {
  __shared__ uint32_t km[kBsKeyWords];
  for (uint32_t i = threadIdx.x; i < kBsKeyWords; i += blockDim.x) km[i] = bs_key_mask(key.rk, i);
  __syncthreads();
  const uint64_t g_first = first_block >> 5, g_last = (first_block + n_blocks - 1) >> 5;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t g = g_first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g <= g_last; g += stride) {
    uint32_t s[128];
    bs_aes_ctr32(km, g << 5, s);
#pragma unroll
    for (uint32_t l = 0; l < 32u; ++l) {
      const uint64_t ctr = (g << 5) + l;
      if (ctr >= first_block && ctr - first_block < n_blocks)
        out[ctr - first_block] = make_uint4(s[l], s[32 + l], s[64 + l], s[96 + l]);
    }
  }
}
#endif

}  // namespace sclgpu
