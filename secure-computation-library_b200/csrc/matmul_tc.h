// matmul_tc.h -- interface between sclgpu.cu and matmul_tc.cu.  Internal to libsclgpu.so.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

#include "field.cuh"

namespace sclgpu {

// per pipeline stage: 128 bytes of an A row (16 Fp61 / 8 Fp127 elements of K); per CTA: 256 limb columns in
// tensor memory (32 Fp61 / 16 Fp127 result columns)
static constexpr uint32_t kMmBTileBytes = 256u * 128u;  // limb image of one (column tile, K chunk) of B

// Fp61 tensor-core path: K even and A 16-byte aligned (a 16-byte cp.async piece holds two elements)
size_t matmul61_image_bytes(uint32_t K, uint32_t N);
size_t matmul127_image_bytes(uint32_t K, uint32_t N);
cudaError_t matmul127_tc_launch(cudaStream_t st, int sm_count, const E127* d_A, uint32_t M, uint32_t K, const E127* d_B, uint32_t N,
                                uint8_t* d_img, E127* d_C);
cudaError_t matmul61_tc_launch(cudaStream_t st, int sm_count, const uint64_t* d_A, uint32_t M, uint32_t K, const uint64_t* d_B,
                               uint32_t N, uint8_t* d_img, uint64_t* d_C);
// warp-specialised form (default): any inner dimension; scratch = limb image of B + pre-tiled A
size_t matmul61_ws_scratch_bytes(uint32_t M, uint32_t K, uint32_t N);
size_t matmul127_ws_scratch_bytes(uint32_t M, uint32_t K, uint32_t N);
cudaError_t matmul61_ws_launch(cudaStream_t st, const uint64_t* d_A, uint32_t M, uint32_t K, const uint64_t* d_B, uint32_t N,
                               uint8_t* d_scratch, uint64_t* d_C);
cudaError_t matmul127_ws_launch(cudaStream_t st, const E127* d_A, uint32_t M, uint32_t K, const E127* d_B, uint32_t N,
                                uint8_t* d_scratch, E127* d_C);
cudaError_t matmul61_generic_launch(cudaStream_t st, int sm_count, const uint64_t* A, uint32_t M, uint32_t K, const uint64_t* B,
                                    uint32_t N, uint64_t* C);
cudaError_t matmul127_generic_launch(cudaStream_t st, int sm_count, const E127* A, uint32_t M, uint32_t K, const E127* B, uint32_t N,
                                     E127* C);

}  // namespace sclgpu
