// share_tc.h -- interface between sclgpu.cu (host logic) and share_tc.cu (the
// tcgen05 Shamir-share kernel).  Internal to libsclgpu.so.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "aes_ctr.cuh"
#include "field.cuh"

namespace sclgpu {

static constexpr int kTcGroups = 3;                   // 128-secret tiles in flight per CTA
static constexpr int kTcThreads = 128 * kTcGroups;    // 4 warps per group
static constexpr uint32_t kTcBmatBytes = 32768;       // 32 parties x 8 limbs rows of K = 128 bytes
static constexpr uint32_t kTcDynSmem = 232448;        // 227 KiB: A tiles | AES tables | B limbs | barriers
// A operand in tensor memory: [1 KiB .. 64 KiB) unused | AES tables | B limbs | barriers.  Requesting exactly
// what is addressed leaves ~4 KiB of the SM's shared memory for a small co-resident CTA (e.g. the HBM-bound
// reconstruction kernel of another stream).
static constexpr uint32_t kTcmDynSmem = (65536u - 1024u) + 131072u + kTcBmatBytes + 128u;
// coefficient-plane variant: B limbs + barriers only; > half an SM so that one CTA owns the tensor memory
static constexpr uint32_t kTcCoeffDynSmem = 120u * 1024u;
static constexpr uint32_t kTcMaxT = 15, kTcMaxParties = 32;        // Fp61:  K = 8(t+1)  <= 128 bytes, 8 limbs per party
static constexpr uint32_t kTcMaxT127 = 7, kTcMaxParties127 = 16;   // Fp127: K = 16(t+1) <= 128 bytes, 16 limbs per party

// Offset of byte (row r = party*8 + limb, column kk = coeff*8 + byte) in the B image:
// K-major rows of 128 bytes, 8-row groups of 1 KiB, 16-byte chunks XOR-swizzled by r%8
// (the canonical SWIZZLE_128B layout tcgen05 reads).
static inline uint32_t tc_bmat_offset(uint32_t r, uint32_t kk) {
  return (r >> 3) * 1024u + (r & 7u) * 128u + (((kk >> 4) ^ (r & 7u)) << 4) + (kk & 15u);
}

// shamirSecretShare + shamirRecoverP in one persistent launch (share_recover.cu).  Lagrange coefficients of the
// reconstruction as three 21-bit limbs each (l0 + l1*2^21 + l2*2^42), zero beyond n: a kernel parameter, read
// from the constant bank.
struct RecBasis61 {
  uint32_t l[32][3];
};
// Destinations of reconstructed secrets when the result is gathered while it is produced: dst[r] is where THIS rank's
// slice starts inside rank r's copy of the gathered vector (own memory or peer memory mapped over NVLink:
// cudaIpcOpenMemHandle / cudaDeviceEnablePeerAccess).  count == 0: the plain output pointer.
struct GatherDst {
  uint64_t* dst[8];
  uint32_t count;
};
cudaError_t share_recover61_prepare();
// d_rec_in == d_shares: reconstruct the sharings produced by this launch (tile by tile, as they are stored);
// otherwise d_rec_in holds another batch of N sharings in the same party-major [n][N] layout.
cudaError_t share_recover61_launch(cudaStream_t st, int sm_count, int variant, const AesKey& key, const RecBasis61& basis,
                                   const uint32_t* d_t0, const void* d_bmat, const void* d_rdimg, uint64_t first_block,
                                   const uint64_t* d_secrets, uint64_t N, uint32_t t, uint32_t n, uint64_t* d_shares,
                                   const uint64_t* d_rec_in, uint64_t* d_rec_out, const GatherDst* gather = nullptr);

cudaError_t share_tc_prepare();
// variant 1: A in shared memory, 3 groups; variants >= 2: A in tensor memory with
// (groups, accumulators per group, columns per MMA pass) as listed in share_tc.cu
int tc_variant_groups(int variant);
cudaError_t share61_tc_launch(int variant, cudaStream_t st, int grid, const AesKey& key, const uint32_t* d_t0, const void* d_bmat,
                              uint64_t first_block, const uint64_t* d_secrets, uint64_t N, uint32_t t, uint32_t n,
                              uint64_t* d_out, uint64_t stride_i, uint64_t stride_j);

cudaError_t share127_tc_launch(int variant, cudaStream_t st, int grid, const AesKey& key, const uint32_t* d_t0,
                               const void* d_bmat, uint64_t first_block, const E127* d_secrets, uint64_t N, uint32_t t,
                               uint32_t n, E127* d_out, uint64_t stride_i, uint64_t stride_j);

// shamirRecoverD / shamirRecoverP on tensor cores: m*BYTES <= 256 (two K tiles of 128 bytes) and
// (n_checks+1)*BYTES <= 128 (two 64-column passes).  Image: K tile kt at kt * kTcRdKTileBytes, inside it the
// canonical layout of up to 128 rows.
static constexpr uint32_t kTcRdKTileBytes = 16384;
static inline uint32_t tc_rd_offset(uint32_t r, uint32_t kk) { return (kk >> 7) * kTcRdKTileBytes + tc_bmat_offset(r, kk & 127u); }
template <class F>
static inline bool recover_d_tc_fits(uint32_t m, uint32_t n_checks) {
  return m >= 1 && m * F::BYTES <= 256u && (n_checks + 1u) * F::BYTES <= 128u;
}
cudaError_t recover_d61_tc_launch(cudaStream_t st, int sm_count, const void* d_bmat, const uint64_t* d_in, uint64_t N,
                                  uint64_t si, uint64_t sj, uint32_t m, uint32_t n_checks, uint64_t* d_out, uint8_t* d_err,
                                  unsigned long long* d_count);
cudaError_t recover_d127_tc_launch(cudaStream_t st, int sm_count, const void* d_bmat, const E127* d_in, uint64_t N,
                                   uint64_t si, uint64_t sj, uint32_t m, uint32_t n_checks, E127* d_out, uint8_t* d_err,
                                   unsigned long long* d_count);

cudaError_t share61_coeffs_tc_launch(cudaStream_t st, int sm_count, const void* d_bmat, const uint64_t* d_coeffs, uint64_t N,
                                     uint32_t t, uint32_t n, uint64_t* d_out, uint64_t stride_i, uint64_t stride_j);
cudaError_t share127_coeffs_tc_launch(cudaStream_t st, int sm_count, const void* d_bmat, const E127* d_coeffs, uint64_t N,
                                      uint32_t t, uint32_t n, E127* d_out, uint64_t stride_i, uint64_t stride_j);

// shamirSecretShare on math::Array<FF, W> with the PRG fused: NW = N*W component polynomials (secrets[j*W + c]),
// sharing j drawing ceil((t+1)*W*BYTES/16) blocks from first_block + j*that.  Fp61: W even.
cudaError_t share61_wide_tc_launch(cudaStream_t st, int sm_count, const AesKey& key, const uint32_t* d_t0, const void* d_bmat,
                                   uint64_t first_block, const uint64_t* d_secrets, uint64_t NW, uint32_t W, uint32_t t, uint32_t n,
                                   uint64_t* d_out, uint64_t stride_i, uint64_t stride_j);
cudaError_t share127_wide_tc_launch(cudaStream_t st, int sm_count, const AesKey& key, const uint32_t* d_t0, const void* d_bmat,
                                    uint64_t first_block, const E127* d_secrets, uint64_t NW, uint32_t W, uint32_t t, uint32_t n,
                                    E127* d_out, uint64_t stride_i, uint64_t stride_j);

}  // namespace sclgpu
