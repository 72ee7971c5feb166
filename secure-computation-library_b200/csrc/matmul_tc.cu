// matmul_tc.cu -- Matrix<Fp>::multiply(Matrix) (include/scl/math/matrix.h:476-495) on the
// 5th-generation tensor cores (both fields), and a generic integer-pipe kernel for the remaining shapes.
//
// Same byte-limb identity as the Shamir kernels (share_tc.cu): with a_{ik} = sum_a a_{ik,a} 2^(8a)
// (the eight bytes of the canonical residue) and C_{kj,a} = b_{kj} 2^(8a) mod p = sum_s C_{kj,a,s} 2^(8s)
//     (A B)_{ij} = sum_s 2^(8s) * acc_{ij,s},    acc_{ij,s} = sum_{k,a} a_{ik,a} * C_{kj,a,s}
// which is a u8 x u8 -> s32 GEMM with inner dimension BYTES*K (exact while BYTES*K * 255^2 < 2^31, i.e. up to
// 32768 bytes = 4096 Fp61 / 2048 Fp127 elements of K per accumulation round).  BYTES = 8 limbs for Fp61, 16 for
// Fp127 (then a, s = 0..15, 2^127 = 1).  B is expanded ONCE into that limb image
// (k_matmul_prep: BYTES^2 bytes per element, stored tile by tile in the canonical 128B-swizzled
// K-major layout tcgen05 reads), A is used as it lies in memory: a row-major row of A is already a
// K-major operand row.
//
// k_matmul_tc<F, RT>: one CTA per (128 RT) x (256 / BYTES) tile of the result.  Per chunk of K (128 bytes of
// an A row): cp.async brings the A chunk (16 KiB, swizzled on the fly) and the matching 32 KiB of
// the limb image into a 4-stage shared-memory ring; one thread issues four
// tcgen05.mma.kind::i8 (M = 128, N = 256, K = 32 bytes) accumulating in 256 TMEM columns;
// tcgen05.commit frees the stage.  Epilogue: each of 128 threads drains its TMEM lane, recombines
// 32 x 8 limbs mod 2^61 - 1 and writes its 32 outputs (one 256-byte row segment).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "field.cuh"
#include "matmul_tc.h"

namespace sclgpu {

static constexpr uint32_t kMmThreads = 256;
static constexpr uint32_t kMmATile = 128u * 128u;                    // 128 rows x 16 elements
// RT = row tiles (of 128) per CTA: 1 -> 4 stages of 48 KiB, 2 -> 3 stages of 64 KiB (the image chunk is shared)
template <int RT> struct MmCfg {
  static constexpr uint32_t kStages = RT == 1 ? 4u : 3u;
  static constexpr uint32_t kStage = RT * kMmATile + kMmBTileBytes;
  static constexpr uint32_t kDynSmem = kStages * kStage + 1024u + 256u;
};
static constexpr uint32_t kMmRoundChunks = 32768u / 128u;             // accumulation round: 32768 bytes of K per A row
static constexpr uint32_t kMmRaster = 8;                              // CTA row tiles per rasterisation band

__device__ __forceinline__ uint32_t mm_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t mm_desc(uint32_t saddr) {  // K-major, SWIZZLE_128B, SBO = 1024 B, sm_100 version
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// D = s32, A = B = u8, K-major, M = 128, N = 256
static constexpr uint32_t kMmIdesc = (2u << 4) | ((256u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void mm_mma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(kMmIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mm_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(mbar) : "memory");
}
__device__ __forceinline__ void mm_wait(uint32_t mbar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mm_cp16(uint32_t dst, const void* src, uint32_t src_bytes) {  // src_bytes 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void mm_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
// sum_s v[s] * 2^(8s) mod p for v[s] < 2^31 (a full 4096-element round): 64-bit pieces, 2^61 = 1
__device__ __forceinline__ uint64_t mm_combine(const uint32_t* v) {
  uint64_t lo = 0, hi = 0;  // value = lo + hi * 2^32, each a sum of four terms < 2^55
#pragma unroll
  for (int s = 0; s < 4; ++s) {
    lo += (uint64_t)v[s] << (8 * s);
    hi += (uint64_t)v[4 + s] << (8 * s);
  }
  // hi * 2^32: (hi mod 2^29) * 2^32 + (hi >> 29) * 2^61
  const uint64_t R = lo + ((hi & 0x1FFFFFFFull) << 32) + (hi >> 29);  // < 2^56 + 2^61 + 2^27
  const uint64_t r = (R & F61::P) + (R >> 61);
  return r >= F61::P ? r - F61::P : r;
}

// sixteen limbs v[s] < 2^31 at 2^(8s) -> canonical residue mod 2^127 - 1.  The four residue classes of s mod 4
// are word aligned (v_r, v_r+4, v_r+8, v_r+12 concatenate into a 128-bit number S_r); X = sum_r S_r << 8r is
// formed in five 32-bit words, folded at bit 127 twice, and p is mapped to 0.
__device__ __forceinline__ E127 mm_combine127(const uint32_t* v) {
  uint32_t x0 = v[0], x1 = v[4], x2 = v[8], x3 = v[12], x4 = 0;
#pragma unroll
  for (int r = 1; r < 4; ++r) {
    const uint32_t a0 = v[r], a1 = v[r + 4], a2 = v[r + 8], a3 = v[r + 12];
    const uint32_t w0 = a0 << (8 * r), w1 = __funnelshift_l(a0, a1, 8 * r), w2 = __funnelshift_l(a1, a2, 8 * r),
                   w3 = __funnelshift_l(a2, a3, 8 * r), w4 = a3 >> (32 - 8 * r);
    asm("add.cc.u32 %0, %0, %5;\n\t"
        "addc.cc.u32 %1, %1, %6;\n\t"
        "addc.cc.u32 %2, %2, %7;\n\t"
        "addc.cc.u32 %3, %3, %8;\n\t"
        "addc.u32 %4, %4, %9;"
        : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3), "+r"(x4)
        : "r"(w0), "r"(w1), "r"(w2), "r"(w3), "r"(w4));
  }
  // X < 2^152: hi = X >> 127 < 2^25
  uint32_t hi = __funnelshift_l(x3, x4, 1);
  x3 &= 0x7FFFFFFFu;
  asm("add.cc.u32 %0, %0, %4;\n\t"
      "addc.cc.u32 %1, %1, 0;\n\t"
      "addc.cc.u32 %2, %2, 0;\n\t"
      "addc.u32 %3, %3, 0;"
      : "+r"(x0), "+r"(x1), "+r"(x2), "+r"(x3)
      : "r"(hi));
  x0 += x3 >> 31;  // < 2^127 + 2^25: if bit 127 is set the rest is < 2^25, no ripple
  x3 &= 0x7FFFFFFFu;
  const bool is_p = (x0 & x1 & x2 & (x3 | 0x80000000u)) == 0xFFFFFFFFu;
  E127 r;
  r.lo = is_p ? 0 : ((uint64_t)x0 | ((uint64_t)x1 << 32));
  r.hi = is_p ? 0 : ((uint64_t)x2 | ((uint64_t)x3 << 32));
  return r;
}

template <class F> struct MmField;
template <> struct MmField<F61> {
  static constexpr uint32_t kKChunk = 16, kNTile = 32;      // 128 / 8, 256 / 8
  static __device__ __forceinline__ uint64_t combine(const uint32_t* v) { return mm_combine(v); }
};
template <> struct MmField<F127> {
  static constexpr uint32_t kKChunk = 8, kNTile = 16;       // 128 / 16, 256 / 16
  static __device__ __forceinline__ E127 combine(const uint32_t* v) { return mm_combine127(v); }
};

// ---- limb image of B: tile (jt, kc) = kNTile columns x kKChunk rows of B -> 256 x 128 bytes, stored at
// ((jt * KC) + kc) * 32 KiB in the canonical swizzled layout.  One CTA per tile; the eight threads q = 0..7 of
// an image row (jl, s) each write its 16-byte chunk q, together the whole 128-byte line.
//   Fp61 : thread (jl 0..31, q): elements B[kc*16 + 2q .. +1][jt*32 + jl], rows s = 0..7
//   Fp127: thread (jl 0..15, q, h 0..1): element B[kc*8 + q][jt*16 + jl], rows s = 8h .. 8h+7
template <class F>
__global__ void __launch_bounds__(256)
k_matmul_prep(const typename F::E* __restrict__ B, uint32_t K, uint32_t N, uint32_t KC, uint8_t* __restrict__ img) {
  const uint32_t kc = blockIdx.x % KC, jt = blockIdx.x / KC;
  const uint32_t q = threadIdx.x & 7u;
  uint8_t* tile = img + (uint64_t)blockIdx.x * kMmBTileBytes;
  if constexpr (F::BYTES == 8) {
    const uint32_t jl = threadIdx.x >> 3;
    const uint32_t j = jt * 32u + jl, k0 = kc * 16u + 2u * q;
    uint64_t c0 = (j < N && k0 < K) ? B[(uint64_t)k0 * N + j] : 0;
    uint64_t c1 = (j < N && k0 + 1 < K) ? B[(uint64_t)(k0 + 1) * N + j] : 0;
    uint64_t ca[16];  // c * 2^(8a), a = 0..7, for both elements
#pragma unroll
    for (int a = 0; a < 8; ++a) {
      ca[a] = c0;
      ca[8 + a] = c1;
      c0 = F61::mul(c0, 256);
      c1 = F61::mul(c1, 256);
    }
#pragma unroll
    for (int s = 0; s < 8; ++s) {
      uint64_t w0 = 0, w1 = 0;  // bytes a = 0..7: byte_s(C_a) of element k0 and of element k0 + 1
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        w0 |= ((ca[a] >> (8 * s)) & 0xFFull) << (8 * a);
        w1 |= ((ca[8 + a] >> (8 * s)) & 0xFFull) << (8 * a);
      }
      const uint32_t r = jl * 8u + s;
      *reinterpret_cast<ulonglong2*>(tile + (r >> 3) * 1024u + (r & 7u) * 128u + ((q ^ (r & 7u)) << 4)) = make_ulonglong2(w0, w1);
    }
  } else {
    const uint32_t jl = (threadIdx.x >> 3) & 15u, h = threadIdx.x >> 7;
    const uint32_t j = jt * 16u + jl, k = kc * 8u + q;
    E127 c = (j < N && k < K) ? B[(uint64_t)k * N + j] : E127{0, 0};
    E127 ca[16];  // c * 2^(8a), a = 0..15
#pragma unroll
    for (int a = 0; a < 16; ++a) {
      ca[a] = c;
      c = F127::mul(c, E127{256, 0});
    }
#pragma unroll
    for (int ss = 0; ss < 8; ++ss) {
      const uint32_t s = h * 8u + ss;  // limb = byte s of the 16-byte residue
      uint64_t w0 = 0, w1 = 0;         // bytes a = 0..7 and a = 8..15 of image row (jl, s)
#pragma unroll
      for (int a = 0; a < 16; ++a) {
        const uint64_t word = (h == 0) ? ca[a].lo : ca[a].hi;
        const uint64_t byte = (word >> (8 * ss)) & 0xFFull;
        if (a < 8) w0 |= byte << (8 * a);
        else w1 |= byte << (8 * (a - 8));
      }
      const uint32_t r = jl * 16u + s;
      *reinterpret_cast<ulonglong2*>(tile + (r >> 3) * 1024u + (r & 7u) * 128u + ((q ^ (r & 7u)) << 4)) = make_ulonglong2(w0, w1);
    }
  }
}

template <class F, int RT>
__global__ void __launch_bounds__(kMmThreads, 1)
k_matmul_tc(const typename F::E* __restrict__ A, uint32_t M, uint32_t K, const uint8_t* __restrict__ img, uint32_t KC,
            uint32_t N, uint32_t NT, typename F::E* __restrict__ C) {
  typedef typename F::E E;
  constexpr uint32_t kMmKChunk = MmField<F>::kKChunk, kMmNTile = MmField<F>::kNTile;
  constexpr uint32_t kPieceElems = 16u / F::BYTES;    // elements per 16-byte cp.async piece
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  constexpr uint32_t kMmStages = MmCfg<RT>::kStages, kMmStage = MmCfg<RT>::kStage;
  constexpr uint32_t kRows = 128u * RT;               // result rows per CTA
  const uint32_t base = (mm_smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t ctl = base + kMmStages * kMmStage;  // empty[stage] mbarriers, done mbarrier, TMEM address
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  // rasterisation: consecutive CTAs walk kMmRaster row tiles of one column tile, then the next column tile,
  // so the ~148 resident CTAs share 16 A row-tiles and ~9 image column-tiles through L2 while they advance
  // along K in step (otherwise the 64-byte-per-element image is re-read from HBM once per row tile)
  const uint32_t MT = (M + kRows - 1u) / kRows;
  const uint32_t per_band = kMmRaster * NT;
  const uint32_t band = blockIdx.x / per_band, in_band = blockIdx.x % per_band;
  const uint32_t band_rows = min(kMmRaster, MT - band * kMmRaster);
  const uint32_t mt = band * kMmRaster + in_band % band_rows, jt = in_band / band_rows;
  if (jt >= NT) return;  // only in the last, shorter band (uniform per CTA: nothing allocated yet)
  const uint32_t m0 = mt * kRows;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ctl + 64u), "n"(256 * RT) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (uint32_t i = 0; i <= kMmStages; ++i)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ctl + 8u * i) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(ctl + 64u) : "memory");
  const uint32_t done_bar = ctl + 8u * kMmStages;

  const uint8_t* img_row = img + (uint64_t)jt * KC * kMmBTileBytes;
  auto load_chunk = [&](uint32_t kc) {
    const uint32_t st = base + (kc % kMmStages) * kMmStage;
    // A: RT tiles of 128 rows x 8 sixteen-byte pieces (row r of the CTA lives in tile r / 128)
#pragma unroll
    for (uint32_t q = tid; q < 1024u * RT; q += kMmThreads) {
      const uint32_t row = q >> 3, piece = q & 7u;
      const uint32_t k = kc * kMmKChunk + piece * kPieceElems;
      const bool in = (m0 + row < M) && (k < K);  // Fp61: K is even, so a piece is inside or outside as a whole
      const E* src = A + (uint64_t)(in ? m0 + row : 0) * K + (in ? k : 0);
      mm_cp16(st + (row >> 3) * 1024u + (row & 7u) * 128u + ((piece ^ (row & 7u)) << 4), src, in ? 16u : 0u);
    }
    // limb image tile: already in shared-memory layout
    const uint8_t* bsrc = img_row + (uint64_t)kc * kMmBTileBytes;
#pragma unroll
    for (uint32_t q = tid; q < kMmBTileBytes / 16u; q += kMmThreads) mm_cp16(st + RT * kMmATile + q * 16u, bsrc + q * 16u, 16u);
  };

  E acc[kMmNTile];
#pragma unroll
  for (uint32_t j = 0; j < kMmNTile; ++j) acc[j] = F::zero();
  uint32_t done_phase = 0;
  // empty[s] completes once per chunk that used stage s (tcgen05.commit after its MMAs), in chunk order:
  // before chunk kc (>= stages) may overwrite its stage, completion number kc/stages - 1 must have happened
  auto wait_stage_free = [&](uint32_t kc) {
    if (kc >= kMmStages) mm_wait(ctl + 8u * (kc % kMmStages), ((kc / kMmStages) - 1u) & 1u);
  };

  for (uint32_t r0 = 0; r0 < KC; r0 += kMmRoundChunks) {
    const uint32_t r1 = min(r0 + kMmRoundChunks, KC);
    // prologue: stages - 1 chunks in flight
    for (uint32_t p = 0; p < kMmStages - 1; ++p) {
      const uint32_t kc = r0 + p;
      if (kc < r1) {
        wait_stage_free(kc);
        load_chunk(kc);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (uint32_t kc = r0; kc < r1; ++kc) {
      const uint32_t nx = kc + kMmStages - 1;
      if (nx < r1) {
        wait_stage_free(nx);
        load_chunk(nx);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group %0;" ::"n"(kMmStages - 1) : "memory");  // this thread's pieces of chunk kc landed
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + (kc % kMmStages) * kMmStage;
#pragma unroll
        for (uint32_t rt = 0; rt < (uint32_t)RT; ++rt)
#pragma unroll
          for (uint32_t ks = 0; ks < 4; ++ks)
            mm_mma(tmem + rt * 256u, mm_desc(st + rt * kMmATile + ks * 32u), mm_desc(st + RT * kMmATile + ks * 32u), (kc > r0) | ks);
        mm_commit(ctl + 8u * (kc % kMmStages));
        if (kc + 1 == r1) mm_commit(done_bar);
      }
    }
    // drain this accumulation round (at most 32768 bytes of K per row: the s32 accumulators are exact)
    mm_wait(done_bar, done_phase);
    done_phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp < 4u * RT) {  // warps 0-3 drain row tile 0, warps 4-7 row tile 1; a warp reads TMEM lanes 32 * (warp % 4) ...
      const uint32_t lane_off = ((warp & 3u) * 32u) << 16;
#pragma unroll
      for (uint32_t g = 0; g < 8; ++g) {
        uint32_t v[32];
        mm_tmem_ld32(tmem + (warp >> 2) * 256u + lane_off + g * 32u, v);
        constexpr uint32_t kPer = 32u / F::BYTES;  // outputs per 32 TMEM columns
#pragma unroll
        for (uint32_t jj = 0; jj < kPer; ++jj)
          acc[g * kPer + jj] = F::add(acc[g * kPer + jj], MmField<F>::combine(v + F::BYTES * jj));
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();  // TMEM drained before the next round overwrites it
  }
  if (warp < 4u * RT) {
    const uint32_t row = m0 + tid;
    if (row < M) {
      E* dst = C + (uint64_t)row * N + (uint64_t)jt * kMmNTile;
#pragma unroll
      for (uint32_t j = 0; j < kMmNTile; ++j)
        if (jt * kMmNTile + j < N) dst[j] = acc[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256 * RT) : "memory");
  }
}

// ============================================================================================================
// Second form: pre-tiled A, bulk async copies, warp-specialised roles (the default; the cp.async form above is
// kept behind SCLGPU_MATMUL_V1 as the measured comparison).
//   k_matmul_prep_a : A (row-major) -> tiles of 128 rows x 128 bytes in the swizzled K-major layout, zero padded to
//                     an even number of row tiles, so that a stage is three contiguous bulk copies (any K, any M).
//   k_matmul_ws     : 10 warps per CTA, one 256 x (256/BYTES) result tile per CTA.
//                     warp 0 (one lane): producer -- cp.async.bulk of two A tiles and one image tile per stage into a
//                                        3-stage ring, completion by mbarrier transaction bytes;
//                     warp 1 (one lane): tcgen05.mma issuer -- 8 MMAs per stage, tcgen05.commit frees the stage;
//                     warps 2-9        : epilogue -- each drains one 32-lane quadrant of one of the two accumulators.
template <class F>
__global__ void __launch_bounds__(256)
k_matmul_prep_a(const typename F::E* __restrict__ A, uint32_t M, uint32_t K, uint32_t KC, uint8_t* __restrict__ tiles) {
  typedef typename F::E E;
  constexpr uint32_t kChunk = MmField<F>::kKChunk, kPiece = 16u / F::BYTES;
  const uint32_t kc = blockIdx.x % KC, mt = blockIdx.x / KC;
  uint8_t* tile = tiles + (uint64_t)blockIdx.x * kMmATile;
#pragma unroll
  for (uint32_t q = threadIdx.x; q < 1024u; q += 256u) {
    const uint32_t row = q >> 3, piece = q & 7u;
    const uint32_t m = mt * 128u + row, k = kc * kChunk + piece * kPiece;
    E e[kPiece];
#pragma unroll
    for (uint32_t i = 0; i < kPiece; ++i) e[i] = (m < M && k + i < K) ? A[(uint64_t)m * K + k + i] : F::zero();
    uint8_t* dst = tile + (row >> 3) * 1024u + (row & 7u) * 128u + ((piece ^ (row & 7u)) << 4);
    if constexpr (F::BYTES == 8) *reinterpret_cast<ulonglong2*>(dst) = make_ulonglong2(e[0], e[1]);
    else *reinterpret_cast<E127*>(dst) = e[0];
  }
}

__device__ __forceinline__ void mm_bulk(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(mbar)
               : "memory");
}

static constexpr uint32_t kWsMmThreads = 320;
static constexpr uint32_t kWsMmStages = 3;
static constexpr uint32_t kWsMmStage = 2u * kMmATile + kMmBTileBytes;  // 64 KiB
static constexpr uint32_t kWsMmDynSmem = kWsMmStages * kWsMmStage + 1024u + 256u;

template <class F>
__global__ void __launch_bounds__(kWsMmThreads, 1)
k_matmul_ws(const uint8_t* __restrict__ a_tiles, uint32_t M, uint32_t MT2, const uint8_t* __restrict__ img, uint32_t KC, uint32_t N,
            uint32_t NT, typename F::E* __restrict__ C) {
  typedef typename F::E E;
  constexpr uint32_t kNTile = MmField<F>::kNTile;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t base = (mm_smem_u32(dyn_smem) + 1023u) & ~1023u;
  const uint32_t ctl = base + kWsMmStages * kWsMmStage;
  const uint32_t bar_full = ctl, bar_empty = ctl + 32u, bar_done = ctl + 64u, bar_drained = ctl + 72u, tmem_slot = ctl + 96u;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;

  const uint32_t per_band = kMmRaster * NT;
  const uint32_t band = blockIdx.x / per_band, in_band = blockIdx.x % per_band;
  const uint32_t band_rows = min(kMmRaster, MT2 - band * kMmRaster);
  const uint32_t mt2 = band * kMmRaster + in_band % band_rows, jt = in_band / band_rows;
  if (jt >= NT) return;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (uint32_t i = 0; i < kWsMmStages; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_full + 8u * i) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_empty + 8u * i) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_done) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(bar_drained) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

  const uint32_t rounds = (KC + kMmRoundChunks - 1u) / kMmRoundChunks;
  if (warp == 0) {
    if (lane == 0) {  // ---------------------------------------------------------------- producer
      const uint8_t* a0 = a_tiles + (uint64_t)(2u * mt2) * KC * kMmATile;
      const uint8_t* a1 = a0 + (uint64_t)KC * kMmATile;
      const uint8_t* bi = img + (uint64_t)jt * KC * kMmBTileBytes;
      for (uint32_t kc = 0; kc < KC; ++kc) {
        const uint32_t s = kc % kWsMmStages;
        if (kc >= kWsMmStages) mm_wait(bar_empty + 8u * s, ((kc / kWsMmStages) - 1u) & 1u);
        const uint32_t st = base + s * kWsMmStage, fb = bar_full + 8u * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(kWsMmStage) : "memory");
        mm_bulk(st, a0 + (uint64_t)kc * kMmATile, kMmATile, fb);
        mm_bulk(st + kMmATile, a1 + (uint64_t)kc * kMmATile, kMmATile, fb);
        mm_bulk(st + 2u * kMmATile, bi + (uint64_t)kc * kMmBTileBytes, kMmBTileBytes, fb);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ---------------------------------------------------------------- MMA issuer
      for (uint32_t kc = 0; kc < KC; ++kc) {
        const uint32_t s = kc % kWsMmStages, r = kc / kMmRoundChunks, first = r * kMmRoundChunks;
        if (kc == first && r > 0) mm_wait(bar_drained, (r - 1u) & 1u);  // the epilogue has read the previous round out of TMEM
        mm_wait(bar_full + 8u * s, (kc / kWsMmStages) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + s * kWsMmStage;
#pragma unroll
        for (uint32_t rt = 0; rt < 2u; ++rt)
#pragma unroll
          for (uint32_t ks = 0; ks < 4u; ++ks)
            mm_mma(tmem + rt * 256u, mm_desc(st + rt * kMmATile + ks * 32u), mm_desc(st + 2u * kMmATile + ks * 32u), (kc > first) | ks);
        mm_commit(bar_empty + 8u * s);
        if (kc + 1u == KC || kc + 1u == first + kMmRoundChunks) mm_commit(bar_done);
      }
    }
  } else {  // ---------------------------------------------------------------------------- epilogue warps 2..9
    const uint32_t e = warp - 2u, rt = e >> 2, quad = warp & 3u;  // a warp may only read TMEM lanes 32 * (warp % 4) ...
    E acc[kNTile];
#pragma unroll
    for (uint32_t j = 0; j < kNTile; ++j) acc[j] = F::zero();
    for (uint32_t r = 0; r < rounds; ++r) {
      mm_wait(bar_done, r & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem + rt * 256u + ((quad * 32u) << 16);
#pragma unroll
      for (uint32_t g = 0; g < 8; ++g) {
        uint32_t v[32];
        mm_tmem_ld32(taddr + g * 32u, v);
        constexpr uint32_t kPer = 32u / F::BYTES;
#pragma unroll
        for (uint32_t jj = 0; jj < kPer; ++jj) acc[g * kPer + jj] = F::add(acc[g * kPer + jj], MmField<F>::combine(v + F::BYTES * jj));
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_drained) : "memory");
    }
    const uint32_t row = mt2 * 256u + rt * 128u + quad * 32u + lane;
    if (row < M) {
      E* dst = C + (uint64_t)row * N + (uint64_t)jt * kNTile;
#pragma unroll
      for (uint32_t j = 0; j < kNTile; ++j)
        if (jt * kNTile + j < N) dst[j] = acc[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
  }
}

// ---- generic kernel (any field, any shape): thread = one output element, lazy accumulation
template <class F>
__global__ void __launch_bounds__(256)
k_matmul_generic(const typename F::E* __restrict__ A, uint32_t M, uint32_t K, const typename F::E* __restrict__ B, uint32_t N,
                 typename F::E* __restrict__ C) {
  const uint64_t total = (uint64_t)M * N;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const uint32_t i = (uint32_t)(idx / N), j = (uint32_t)(idx % N);
    typename F::Acc acc = F::acc_zero();
    int terms = 0;
    for (uint32_t k = 0; k < K; ++k) {
      F::mac(acc, A[(uint64_t)i * K + k], B[(uint64_t)k * N + j]);
      if (++terms == F::ACC_TERMS - 1) {
        F::acc_fold(acc);
        terms = 0;
      }
    }
    C[idx] = F::acc_reduce(acc);
  }
}

template <class F>
static size_t image_bytes_t(uint32_t K, uint32_t N) {
  const uint64_t KC = (K + MmField<F>::kKChunk - 1) / MmField<F>::kKChunk, NT = (N + MmField<F>::kNTile - 1) / MmField<F>::kNTile;
  return (size_t)(KC * NT * kMmBTileBytes);
}
size_t matmul61_image_bytes(uint32_t K, uint32_t N) { return image_bytes_t<F61>(K, N); }
size_t matmul127_image_bytes(uint32_t K, uint32_t N) { return image_bytes_t<F127>(K, N); }

template <class F>
static cudaError_t matmul_tc_launch_t(cudaStream_t st, const typename F::E* d_A, uint32_t M, uint32_t K, const typename F::E* d_B,
                                      uint32_t N, uint8_t* d_img, typename F::E* d_C) {
  const uint32_t KC = (K + MmField<F>::kKChunk - 1) / MmField<F>::kKChunk, NT = (N + MmField<F>::kNTile - 1) / MmField<F>::kNTile;
  k_matmul_prep<F><<<KC * NT, 256, 0, st>>>(d_B, K, N, KC, d_img);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  if (M > 128) {  // two row tiles per CTA: every image chunk feeds twice the MMAs
    e = cudaFuncSetAttribute(k_matmul_tc<F, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MmCfg<2>::kDynSmem);
    if (e != cudaSuccess) return e;
    const uint32_t MT = (M + 255) / 256, bands = (MT + kMmRaster - 1) / kMmRaster;
    k_matmul_tc<F, 2><<<bands * kMmRaster * NT, kMmThreads, MmCfg<2>::kDynSmem, st>>>(d_A, M, K, d_img, KC, N, NT, d_C);
  } else {
    e = cudaFuncSetAttribute(k_matmul_tc<F, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MmCfg<1>::kDynSmem);
    if (e != cudaSuccess) return e;
    const uint32_t MT = (M + 127) / 128, bands = (MT + kMmRaster - 1) / kMmRaster;
    k_matmul_tc<F, 1><<<bands * kMmRaster * NT, kMmThreads, MmCfg<1>::kDynSmem, st>>>(d_A, M, K, d_img, KC, N, NT, d_C);
  }
  return cudaGetLastError();
}
// scratch layout of the warp-specialised form: [limb image of B][tiles of A, padded to an even number of row tiles]
template <class F>
static size_t ws_scratch_bytes_t(uint32_t M, uint32_t K, uint32_t N) {
  const uint64_t KC = (K + MmField<F>::kKChunk - 1) / MmField<F>::kKChunk, MT2 = (M + 255) / 256;
  return image_bytes_t<F>(K, N) + (size_t)(2 * MT2 * KC * kMmATile);
}
size_t matmul61_ws_scratch_bytes(uint32_t M, uint32_t K, uint32_t N) { return ws_scratch_bytes_t<F61>(M, K, N); }
size_t matmul127_ws_scratch_bytes(uint32_t M, uint32_t K, uint32_t N) { return ws_scratch_bytes_t<F127>(M, K, N); }

template <class F>
static cudaError_t matmul_ws_launch_t(cudaStream_t st, const typename F::E* d_A, uint32_t M, uint32_t K, const typename F::E* d_B,
                                      uint32_t N, uint8_t* d_scratch, typename F::E* d_C) {
  const uint32_t KC = (K + MmField<F>::kKChunk - 1) / MmField<F>::kKChunk, NT = (N + MmField<F>::kNTile - 1) / MmField<F>::kNTile;
  const uint32_t MT2 = (M + 255) / 256;
  uint8_t* d_img = d_scratch;
  uint8_t* d_at = d_scratch + image_bytes_t<F>(K, N);
  k_matmul_prep<F><<<KC * NT, 256, 0, st>>>(d_B, K, N, KC, d_img);
  k_matmul_prep_a<F><<<2 * MT2 * KC, 256, 0, st>>>(d_A, M, K, KC, d_at);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_matmul_ws<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWsMmDynSmem);
  if (e != cudaSuccess) return e;
  const uint32_t bands = (MT2 + kMmRaster - 1) / kMmRaster;
  k_matmul_ws<F><<<bands * kMmRaster * NT, kWsMmThreads, kWsMmDynSmem, st>>>(d_at, M, MT2, d_img, KC, N, NT, d_C);
  return cudaGetLastError();
}
cudaError_t matmul61_ws_launch(cudaStream_t st, const uint64_t* d_A, uint32_t M, uint32_t K, const uint64_t* d_B, uint32_t N,
                               uint8_t* d_scratch, uint64_t* d_C) {
  return matmul_ws_launch_t<F61>(st, d_A, M, K, d_B, N, d_scratch, d_C);
}
cudaError_t matmul127_ws_launch(cudaStream_t st, const E127* d_A, uint32_t M, uint32_t K, const E127* d_B, uint32_t N,
                                uint8_t* d_scratch, E127* d_C) {
  return matmul_ws_launch_t<F127>(st, d_A, M, K, d_B, N, d_scratch, d_C);
}

cudaError_t matmul61_tc_launch(cudaStream_t st, int, const uint64_t* d_A, uint32_t M, uint32_t K, const uint64_t* d_B, uint32_t N,
                               uint8_t* d_img, uint64_t* d_C) {
  return matmul_tc_launch_t<F61>(st, d_A, M, K, d_B, N, d_img, d_C);
}
cudaError_t matmul127_tc_launch(cudaStream_t st, int, const E127* d_A, uint32_t M, uint32_t K, const E127* d_B, uint32_t N,
                                uint8_t* d_img, E127* d_C) {
  return matmul_tc_launch_t<F127>(st, d_A, M, K, d_B, N, d_img, d_C);
}

cudaError_t matmul61_generic_launch(cudaStream_t st, int sm_count, const uint64_t* A, uint32_t M, uint32_t K, const uint64_t* B,
                                    uint32_t N, uint64_t* C) {
  const uint64_t total = (uint64_t)M * N;
  k_matmul_generic<F61><<<(int)std::min<uint64_t>((total + 255) / 256, (uint64_t)sm_count * 8), 256, 0, st>>>(A, M, K, B, N, C);
  return cudaGetLastError();
}
cudaError_t matmul127_generic_launch(cudaStream_t st, int sm_count, const E127* A, uint32_t M, uint32_t K, const E127* B, uint32_t N,
                                     E127* C) {
  const uint64_t total = (uint64_t)M * N;
  k_matmul_generic<F127><<<(int)std::min<uint64_t>((total + 255) / 256, (uint64_t)sm_count * 8), 256, 0, st>>>(A, M, K, B, N, C);
  return cudaGetLastError();
}

}  // namespace sclgpu
