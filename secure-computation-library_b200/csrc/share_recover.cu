// share_recover.cu -- the C2 step in ONE persistent launch: shamirSecretShare of a batch (PRG fused, tcgen05)
// with shamirRecoverP running beside it on the same SMs.
//
// Reference path replaced: N x { ss::shamirSecretShare (include/scl/ss/shamir.h:52-68) on one PRG,
// ss::shamirRecoverP (shamir.h:82-104: Lagrange basis at 0 through nodes 1..n, lagrange.h:55-71, then
// innerProd, vector.h:45-52) } -- what a dealer's "share, send, collect, reconstruct" round does per secret.
//
// Why one kernel.  Measured on B200 (profiles/r01k_*): the share kernel k_share_tcm<F61,5,1,64> is bound by the
// shared-memory data pipe (T-table AES-CTR, 88 %) and the ALU pipe (79 %), DRAM 20 %; the reconstruction kernel
// k_recover61_pm is bound by HBM (DRAM 76 %), LSU 8 %, its multiplications on the FMA pipe the share kernel
// leaves idle.  Run back to back they take 10.7 + 2.9 ms; the resources they saturate are disjoint.  Here the
// five share groups of k_share_tcm run unchanged, and RWARPS extra warps of the same CTA reconstruct (two secrets
// per thread through 128-bit loads, eight planes in flight per thread):
//   * dependent mode (rec_in == share output): the planes of a 128-secret tile as soon as its share group has
//     stored them (per-group tile counters in shared memory; a CTA only ever consumes its own tiles, so no CTA
//     waits for another and the launch cannot deadlock whatever the residency) -- the reads are L2 hits;
//   * independent mode: the planes of ANOTHER batch (e.g. the one shared by the previous launch) -- HBM reads
//     under the LSU-bound share work.
// The reconstruction arithmetic is k_recover61_pm's (kernels.cuh): Lagrange coefficient in three 21-bit limbs,
// share in two 32-bit words, six IMAD.WIDE per term accumulating in 64 bits without carries (n <= 32 terms of
// < 2^53), one recombination per secret by 61-bit rotations.  The limbs come from the constant bank (kernel
// parameter), so a term costs no shared-memory access and no register for the coefficient.
// Measured alternatives (B200, 2^26 secrets, n=32, t=15; share alone 10.75 ms, reconstruction alone 2.86 ms):
//   * planes staged by cp.async.bulk into a 3 x 16 KiB shared-memory ring + 4 consumer warps: 13.4 ms -- the bulk
//     writes and the LDS re-reads both go through the shared-memory pipe the AES already saturates (the loader
//     alone cost +1.0 ms, LDS + arithmetic +1.4 ms);
//   * the two kernels on two streams: 12.8 ms;
//   * homogeneous groups (round 2): no reconstruction group -- every group shares a tile, then reconstructs one on the
//     tensor core inside its own 96 tensor-memory columns, which lets a FIFTH group fit: 13.07 ms (planes loaded 16 at a
//     time), 16.6 ms (all 32 at once: spills), 13.5 ms with four groups, against 12.28 ms for <4,4,tc>: the plane loads
//     cannot be requested a tile ahead (no registers beside the AES state), so each group waits for them in line.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>

#include "aes_ctr.cuh"
#include "field.cuh"
#include "share_tc.h"
#include "tc_common.cuh"

namespace sclgpu {

static constexpr uint32_t kSrDynSmem = kTcmDynSmem;

namespace {

__device__ __forceinline__ uint64_t rot61(uint64_t acc, int s) {  // acc * 2^s mod p (semi-canonical in, < 2^61 out)
  uint64_t x = (acc & F61::P) + (acc >> 61);
  x = x >= F61::P ? x - F61::P : x;
  return s == 0 ? x : (((x << s) & F61::P) | (x >> (61 - s)));
}

}  // namespace

// GROUPS share groups of 4 warps + RWARPS reconstruction warps.  Fp61, t <= 15, n <= 32, party-major planes:
// share (j, i) at out[i * N + j].  rec_in: planes [n][N] to reconstruct (== out for the dependent mode).
// TCREC: the reconstruction as a byte-limb product on the tensor core (k_recover_d_tc's formulation: the 8n share bytes
// of a secret as one A row in tensor memory, the limbs of lambda_i * 2^(8a) as B, g_rdimg) instead of IMAD chains.
template <int GROUPS, int RWARPS, bool TCREC>
__global__ void __launch_bounds__(128 * GROUPS + 32 * RWARPS, 1)
k_share_recover61(const __grid_constant__ AesKey key, const __grid_constant__ RecBasis61 basis,
                  const uint32_t* __restrict__ g_t0, const uint4* __restrict__ g_bmat, const uint4* __restrict__ g_rdimg,
                  uint64_t first_block,
                  const uint64_t* __restrict__ secrets, uint64_t N, uint32_t t, uint32_t n, uint64_t* out,
                  const uint64_t* rec_in, uint64_t* __restrict__ rec_out, uint32_t dependent,
                  const __grid_constant__ GatherDst gather) {
  // warp roles: [0, RWARPS) reconstruction, then GROUPS x 4 share warps
  constexpr uint32_t kRecWarps = RWARPS, kRecThreads = 32 * kRecWarps;
  static_assert(RWARPS % 4 == 0, "share warp w must sit on tensor-memory lane quarter w % 4");
  constexpr uint32_t kThreads = kRecThreads + 128 * GROUPS;
  constexpr uint32_t PCOLS = 64, kColsPerGroup = 32u + PCOLS, kPassParties = 8, kLdParties = 4;
  constexpr uint32_t kRecACols = 64, kRecDCols = 16;  // TCREC: 256 share bytes per lane, one 16-column accumulator
  static_assert(GROUPS * kColsPerGroup + (TCREC ? kRecACols + kRecDCols : 0u) <= 512, "tensor memory has 512 columns");
  static_assert(!TCREC || RWARPS == 4, "the tensor-core reconstruction is one group of four warps");
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  const uint32_t dyn = smem_u32(dyn_smem);
  const uint32_t tbase = aes_table_base(dyn_smem);
  const uint32_t b_base = tbase + kAesTableBytes;
  // control block: GROUPS mbarriers | GROUPS tile counters | TMEM base address
  const uint32_t ctl = b_base + kTcBmatBytes;
  const uint32_t done0 = ctl + 64u;            // u32 per group: warps of that group that have stored a tile
  const uint32_t tmem_slot = ctl + 96u, mbar_rec = ctl + 48u;
  // TCREC: the reconstruction's B image (two K tiles of 16 rows x 128 bytes) in the free shared memory below the tables
  const uint32_t brec = (dyn + 1023u) & ~1023u;
  if (ctl + 128u > dyn + kSrDynSmem || brec + 4096u > tbase) __trap();

  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u;
  aes_fill_tables(tbase, g_t0);
  for (uint32_t e = tid; e < kTcBmatBytes / 16; e += kThreads) {
    const uint4 w = __ldg(g_bmat + e);
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(b_base + e * 16u), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
  }
  if constexpr (TCREC) {
    for (uint32_t e = tid; e < 256u; e += kThreads) {  // 2 x 2 KiB: rows 0..15 of each K tile of the image
      const uint4 w = __ldg(g_rdimg + (e >> 7) * (kTcRdKTileBytes / 16u) + (e & 127u));
      asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(brec + e * 16u), "r"(w.x), "r"(w.y), "r"(w.z), "r"(w.w) : "memory");
    }
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (uint32_t i = 0; i < GROUPS; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ctl + 8u * i) : "memory");
      asm volatile("st.shared.u32 [%0], %1;" ::"r"(done0 + 4u * i), "r"(0u) : "memory");
    }
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_rec) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  uint32_t tmem;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem) : "r"(tmem_slot) : "memory");

  const uint64_t tiles = (N + 127u) / 128u;
  const uint64_t tile_step = (uint64_t)gridDim.x * GROUPS;

  if (tid >= kRecThreads) {
    // ================================================================ share groups (k_share_tcm<F61, GROUPS, 1, 64>)
    uint32_t lanebase = tbase + lane * 4u;
    asm volatile("" : "+r"(lanebase)::"memory");
    const uint32_t g = (tid - kRecThreads) >> 7, gt = tid & 127u;
    const uint32_t a_tm = tmem + g * kColsPerGroup;  // 32 columns: 128 coefficient bytes per lane
    const uint32_t acc0 = a_tm + 32u;
    const uint32_t lane_off = ((warp & 3u) * 32u) << 16;
    const uint32_t mbar = ctl + 8u * g;
    uint32_t ph = 0;
    const uint32_t nblk = ((t + 1u) * 8u + 15u) / 16u;    // keystream blocks per secret (prg.cc:129-133)
    const uint32_t ksteps = ((t + 1u) * 8u + 31u) / 32u;  // K = 32 bytes per MMA
    const uint32_t npass = (n + kPassParties - 1u) / kPassParties;

    auto issue_pass = [&](uint32_t p) {
      for (uint32_t ks = 0; ks < ksteps; ++ks)
        tc_mma_ts(acc0, a_tm + ks * 8u, tc_desc(b_base + p * (PCOLS * 128u) + ks * 32u), tc_idesc(PCOLS), ks);
      tc_commit(mbar);
    };
    auto emit = [&](const uint32_t (&v)[32], uint64_t* dst, uint32_t first_party) {
      if (first_party + kLdParties <= n) {
#pragma unroll
        for (uint32_t ii = 0; ii < kLdParties; ++ii) dst[(uint64_t)ii * N] = tc_combine(v + 8 * ii);
        return;
      }
#pragma unroll
      for (uint32_t ii = 0; ii < kLdParties; ++ii)
        if (first_party + ii < n) dst[(uint64_t)ii * N] = tc_combine(v + 8 * ii);
    };

    uint32_t produced = 0, published = 0;  // tiles stored / announced by this warp
    auto publish = [&]() {
      // every lane's stores are ordered before the warp's increment of the group counter
      __threadfence_block();
      __syncwarp();
      if (lane == 0) asm volatile("red.release.cta.shared::cta.add.u32 [%0], %1;" ::"r"(done0 + 4u * g), "r"(1u) : "memory");
      ++published;
    };
    // Tiles of this group.  IMAD-reconstruction forms: strided over the grid (the reconstruction warps follow the same
    // order).  Tensor-core form: a CONTIGUOUS range, so that a thread's consecutive tiles are consecutive 256-counter
    // groups of the PRG and the group state can be carried over (prg_group_cached: 5 lookups per tile instead of 27).
    uint64_t tile_begin, tile_end, tile_inc;
    if constexpr (TCREC) {
      const uint64_t gi = (uint64_t)blockIdx.x * GROUPS + g, n_groups = (uint64_t)gridDim.x * GROUPS;
      tile_begin = gi * tiles / n_groups;
      tile_end = (gi + 1u) * tiles / n_groups;
      tile_inc = 1;
    } else {
      tile_begin = (uint64_t)blockIdx.x * GROUPS + g;
      tile_end = tiles;
      tile_inc = tile_step;
    }
    PrgGroupCache gcache;
    gcache.tag = ~0ull;
    for (uint64_t tile = tile_begin; tile < tile_end; tile += tile_inc) {
      const uint64_t j = tile * 128u + gt;
      const bool valid = j < N;
      const uint64_t jj = valid ? j : N - 1;  // tail lanes recompute the last secret (never stored)
      const uint32_t a_lane = a_tm + lane_off;
      const uint64_t sec = secrets[jj];
      const uint32_t s0 = (uint32_t)sec, s1 = (uint32_t)(sec >> 32);  // coefficient 0 (shamir.h:56-57)
      const uint64_t ctr0 = first_block + jj * nblk;
      if (t == 0) {  // one block is consumed, none of it is used
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %3};" ::"r"(a_lane), "r"(s0), "r"(s1), "r"(0u) : "memory");
      } else {
        PrgGroup grp;
        if constexpr (TCREC) prg_group_cached(key, lanebase, ctr0, grp, gcache);
        else prg_group(key, lanebase, ctr0, grp);
        // the block of this secret (if any; nblk <= 8) that opens the next 256-counter group
        const uint32_t cross = 256u - ((uint32_t)ctr0 & 255u), c_lo = (uint32_t)ctr0;
        if ((nblk & 1u) == 0u && __all_sync(0xffffffffu, cross >= nblk)) {
          // TWO blocks per iteration (no group crossing inside this warp's secrets: the usual case): two independent
          // lookup chains for the scheduler to interleave.  With one block at a time a warp stalls at every round
          // boundary (XOR tree -> address -> lookup); measured 12.33 -> 11.89 ms for the step on one box.  Four blocks
          // per iteration: 12.20 ms (the 96-register budget); three share groups (128 registers) with four: 12.92 ms.
#pragma unroll 1
          for (uint32_t b = 0; b < nblk; b += 2u) {
            uint32_t o0 = s0, o1 = s1, o2, o3, q0, q1, q2, q3;  // block 0: coefficient 0 is the secret
            prg_block_grouped(key, lanebase, grp, c_lo + b, o0, o1, o2, o3, b != 0);
            prg_block_grouped(key, lanebase, grp, c_lo + b + 1u, q0, q1, q2, q3);
            __syncwarp();
            // keystream blocks b, b + 1 = K bytes [16b, 16b + 32) of the row = TMEM columns 4b..4b+7 of this lane
            asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(a_lane + 4u * b),
                         "r"(o0), "r"(o1), "r"(o2), "r"(o3), "r"(q0), "r"(q1), "r"(q2), "r"(q3) : "memory");
          }
        } else
#pragma unroll 1
        for (uint32_t b = 0; b < nblk; ++b) {
          if (b == cross) {  // at most once per secret
            if constexpr (TCREC) prg_group_cached(key, lanebase, ctr0 + b, grp, gcache);
            else prg_group(key, lanebase, ctr0 + b, grp);
          }
          uint32_t o0 = s0, o1 = s1, o2, o3;  // block 0: coefficient 0 is the secret, the keystream words are not computed
          prg_block_grouped(key, lanebase, grp, c_lo + b, o0, o1, o2, o3, b != 0);
          __syncwarp();
          // keystream block b = K bytes [16b, 16b+16) of the row = TMEM columns 4b..4b+3 of this lane
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_lane + 4u * b), "r"(o0), "r"(o1), "r"(o2), "r"(o3) : "memory");
        }
      }
      // the PREVIOUS tile of this warp is published here, one AES phase after its stores were issued: by now they
      // have drained, so the fence does not wait (right after the stores it cost ~1 us per tile and warp)
      if (dependent && published < produced) publish();
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      group_sync(g);
      if (gt == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        issue_pass(0);
      }
      for (uint32_t p = 0; p < npass; ++p) {
        mbar_wait(mbar, ph);
        ph ^= 1u;
        __syncwarp();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t acc = acc0 + lane_off;
        uint64_t* dst = out + j + (uint64_t)(p * kPassParties) * N;
        uint32_t v[32];
        tmem_ld32(acc, v);
        if (valid) emit(v, dst, p * kPassParties);
        tmem_ld32(acc + 32u, v);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        group_sync(g);  // accumulator drained by every warp of the group: it may be overwritten
        if (gt == 0 && p + 1u < npass) {
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          issue_pass(p + 1u);
        }
        if (valid) emit(v, dst + (uint64_t)kLdParties * N, p * kPassParties + kLdParties);
      }
      ++produced;
    }
    if (dependent && published < produced) publish();
  } else if constexpr (TCREC) {
    // ================================================================ reconstruction group on the tensor core
    // Thread = secret = tensor-memory lane.  Per 128-secret tile: the n shares of the secret go into the lane's A row
    // (two shares per tcgen05.st; the loads were requested one tile ahead, under the previous tile's MMAs and epilogue),
    // one elected thread issues ceil(8n / 32) MMAs (M = 128, N = 16, K = 32) against the limb
    // image of the Lagrange coefficients, and the 8 limbs of the secret come back with one tcgen05.ld.  Pipelined mode
    // only (the planes of another batch): the prefetch would run ahead of a tile this launch has yet to store.
    const uint32_t a_rec = tmem + GROUPS * kColsPerGroup, d_rec = a_rec + kRecACols;
    const uint32_t lane_off = ((warp & 3u) * 32u) << 16;
    const uint32_t ksteps = (n * 8u + 31u) / 32u;
    uint32_t ph = 0;
    // the next tile's shares, as the 32-bit words tcgen05.st takes (a load can land in the registers the store reads)
    uint32_t w[64];
    auto load_tile = [&](const uint64_t* src) {
#pragma unroll
      for (uint32_t u = 0; u < 32u; ++u)
        if (u < n)
          asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0, %1}, [%2];" : "=r"(w[2 * u]), "=r"(w[2 * u + 1]) : "l"(src + (uint64_t)u * N));
    };
    // this CTA's tiles: the contiguous ranges of its share groups, i.e. one contiguous range
    const uint64_t n_groups = (uint64_t)gridDim.x * GROUPS;
    const uint64_t cta_lo = (uint64_t)blockIdx.x * GROUPS * tiles / n_groups;
    const uint64_t cta_hi = ((uint64_t)blockIdx.x + 1u) * GROUPS * tiles / n_groups;
    if (cta_lo < cta_hi) {
      const uint64_t j0 = cta_lo * 128u + tid;
      load_tile(rec_in + (j0 < N ? j0 : N - 1));
    }
    for (uint64_t tile = cta_lo; tile < cta_hi; ++tile) {
      const uint64_t j = tile * 128u + tid;
#pragma unroll
      for (uint32_t c = 0; c < 16u; ++c)
        if (2u * c < n)  // CTA-uniform; an odd n leaves a stale word in the last pair: its B rows are zero
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_rec + lane_off + 4u * c),
                       "r"(w[4 * c]), "r"(w[4 * c + 1]), "r"(w[4 * c + 2]), "r"(w[4 * c + 3])
                       : "memory");
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      // the stores have read their registers: the next tile's planes are requested now, in flight under this tile's
      // barrier, MMAs and epilogue; the tile after that is pulled into L2 (two 128-byte lines per thread)
      if (tile + 1u < cta_hi) {
        const uint64_t jn = (tile + 1u) * 128u + tid;
        load_tile(rec_in + (jn < N ? jn : N - 1));
      }
      if (tile + 2u < cta_hi) {
#pragma unroll
        for (uint32_t h = 0; h < 2u; ++h) {
          const uint32_t line = tid + 128u * h;  // 32 planes x 8 lines of 16 secrets
          const uint64_t e = (tile + 2u) * 128u + (line & 7u) * 16u;
          if ((line >> 3) < n && e < N)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(rec_in + (uint64_t)(line >> 3) * N + e));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"((uint32_t)GROUPS + 1u) : "memory");
      if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (uint32_t ks = 0; ks < ksteps; ++ks)
          tc_mma_ts(d_rec, a_rec + ks * 8u, tc_desc(brec + (ks >> 2) * 2048u + (ks & 3u) * 32u), tc_idesc(kRecDCols), ks);
        tc_commit(mbar_rec);
      }
      mbar_wait(mbar_rec, ph);
      ph ^= 1u;
      __syncwarp();
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t v[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(d_rec + lane_off)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      const uint64_t r = tc_combine24(v);
      if (j < N) {
        if (gather.count == 0) {
          rec_out[j] = r;
        } else {
#pragma unroll
          for (uint32_t g = 0; g < 8u; ++g)
            if (g < gather.count) gather.dst[g][j] = r;
        }
      }
    }
  } else {
    // ================================================================ reconstruction warps (shamirRecoverP)
    // A 128-secret tile is reconstructed by a pair of warps, two secrets per thread through 128-bit loads, eight
    // planes requested before the first product (k_recover61_pm's schedule).  Pair p of the kRecWarps / 2 pairs takes
    // the CTA's tiles p, p + pairs, ... in the order the share groups produce them: tile q = k * GROUPS + g.
    constexpr uint32_t kPairs = kRecWarps / 2;
    const uint32_t pair = warp >> 1, rt = tid & 63u;  // thread rt of the pair: secrets 2rt, 2rt+1 of the tile
    for (uint32_t q = pair;; q += kPairs) {
      const uint32_t k = q / GROUPS, g = q - k * GROUPS;
      const uint64_t tile = (uint64_t)blockIdx.x * GROUPS + g + (uint64_t)k * tile_step;
      if ((uint64_t)blockIdx.x * GROUPS + (uint64_t)k * tile_step >= tiles) break;
      if (tile >= tiles) continue;
      if (dependent) {  // the tile exists once all four warps of its share group have stored it
        if (lane == 0) {
          const uint32_t want = 4u * (k + 1u);
          for (;;) {
            uint32_t have;
            asm volatile("ld.acquire.cta.shared::cta.u32 %0, [%1];" : "=r"(have) : "r"(done0 + 4u * g) : "memory");
            if (have >= want) break;
            __nanosleep(400);
          }
        }
        __syncwarp();
      }
      const uint64_t j = tile * 128u + 2u * rt;
      const uint64_t* src = rec_in + (j < N ? j : N - 2);  // N is even: a thread's pair is inside or outside
      uint64_t acc[2][6];
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int c = 0; c < 6; ++c) acc[h][c] = 0;
#pragma unroll
      for (uint32_t i0 = 0; i0 < 32u; i0 += 8u) {
        if (i0 < n) {  // CTA-uniform
          uint64_t v[8][2];
#pragma unroll
          for (uint32_t u = 0; u < 8u; ++u) {
            const uint64_t* p = src + (uint64_t)(i0 + u < n ? i0 + u : n - 1u) * N;  // beyond n: any plane, coefficient 0
            if (dependent) {  // written by this CTA during this launch: coherent load (served by L2)
              asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(v[u][0]), "=l"(v[u][1]) : "l"(p));
            } else {
              asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];" : "=l"(v[u][0]), "=l"(v[u][1]) : "l"(p));
            }
          }
#pragma unroll
          for (uint32_t u = 0; u < 8u; ++u) {
            const uint32_t l0 = basis.l[i0 + u][0], l1 = basis.l[i0 + u][1], l2 = basis.l[i0 + u][2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const uint32_t w0 = (uint32_t)v[u][h], w1 = (uint32_t)(v[u][h] >> 32);
              asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[h][0]) : "r"(w0), "r"(l0));
              asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[h][1]) : "r"(w0), "r"(l1));
              asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[h][2]) : "r"(w0), "r"(l2));
              asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[h][3]) : "r"(w1), "r"(l0));
              asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[h][4]) : "r"(w1), "r"(l1));
              asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[h][5]) : "r"(w1), "r"(l2));
            }
          }
        }
      }
      uint64_t r[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        // weights 2^0, 2^21, 2^42, 2^32, 2^53, 2^74 = 2^13
        const uint64_t s = rot61(acc[h][0], 0) + rot61(acc[h][1], 21) + rot61(acc[h][2], 42) + rot61(acc[h][3], 32) +
                           rot61(acc[h][4], 53) + rot61(acc[h][5], 13);  // < 6 * 2^61 < 2^64
        r[h] = F61::from_raw(s);
      }
      if (j < N) {
        if (gather.count == 0) {
          *reinterpret_cast<ulonglong2*>(rec_out + j) = make_ulonglong2(r[0], r[1]);
        } else {
          // all-gather fused in: the pair goes to every rank's copy of the gathered vector (posted stores over NVLink
          // peer memory), hidden under the share groups' work
#pragma unroll
          for (uint32_t g = 0; g < 8u; ++g)
            if (g < gather.count) *reinterpret_cast<ulonglong2*>(gather.dst[g] + j) = make_ulonglong2(r[0], r[1]);
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

static constexpr int kSrGroups = 5;

cudaError_t share_recover61_prepare() {
  cudaError_t e = cudaSuccess;
#define SCLGPU_SR_ATTR(G, RW, TC) \
  if (e == cudaSuccess) e = cudaFuncSetAttribute(k_share_recover61<G, RW, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSrDynSmem);
  SCLGPU_SR_ATTR(kSrGroups, 4, false)
  SCLGPU_SR_ATTR(kSrGroups, 8, false)
  SCLGPU_SR_ATTR(4, 4, false)
  SCLGPU_SR_ATTR(4, 8, false)
  SCLGPU_SR_ATTR(4, 4, true)
#undef SCLGPU_SR_ATTR
  return e;
}

template <int G, int RW, bool TC>
static void sr_launch(cudaStream_t st, int sm_count, const AesKey& key, const RecBasis61& basis, const uint32_t* d_t0,
                      const uint4* bm, const uint4* rdimg, uint64_t first_block, const uint64_t* d_secrets, uint64_t N, uint32_t t,
                      uint32_t n, uint64_t* d_shares, const uint64_t* d_rec_in, uint64_t* d_rec_out, uint32_t dependent,
                      const GatherDst& gd) {
  const uint64_t tiles = (N + 127) / 128;
  const int grid = (int)std::min<uint64_t>((tiles + G - 1) / G, (uint64_t)sm_count);
  k_share_recover61<G, RW, TC><<<grid, 128 * G + 32 * RW, kSrDynSmem, st>>>(key, basis, d_t0, bm, rdimg, first_block, d_secrets, N, t, n,
                                                                            d_shares, d_rec_in, d_rec_out, dependent, gd);
}

// variant: 4 (default) / 8 reconstruction warps beside five share groups; 104 / 108 the same beside four share groups;
// 204 four share groups + the tensor-core reconstruction group (needs d_rdimg and the pipelined mode)
cudaError_t share_recover61_launch(cudaStream_t st, int sm_count, int variant, const AesKey& key, const RecBasis61& basis,
                                   const uint32_t* d_t0, const void* d_bmat, const void* d_rdimg, uint64_t first_block,
                                   const uint64_t* d_secrets, uint64_t N, uint32_t t, uint32_t n, uint64_t* d_shares,
                                   const uint64_t* d_rec_in, uint64_t* d_rec_out, const GatherDst* gather) {
  GatherDst gd{};
  if (gather) gd = *gather;
  const uint32_t dependent = (d_rec_in == d_shares) ? 1u : 0u;
  const uint4* bm = reinterpret_cast<const uint4*>(d_bmat);
  const uint4* rd = reinterpret_cast<const uint4*>(d_rdimg);
  if (variant == 204 && (dependent || rd == nullptr)) variant = 104;  // the tensor-core form needs another batch
#define SCLGPU_SR_ARGS st, sm_count, key, basis, d_t0, bm, rd, first_block, d_secrets, N, t, n, d_shares, d_rec_in, d_rec_out, dependent, gd
  switch (variant) {
    case 8: sr_launch<kSrGroups, 8, false>(SCLGPU_SR_ARGS); break;
    case 104: sr_launch<4, 4, false>(SCLGPU_SR_ARGS); break;
    case 108: sr_launch<4, 8, false>(SCLGPU_SR_ARGS); break;
    case 204: sr_launch<4, 4, true>(SCLGPU_SR_ARGS); break;
    default: sr_launch<kSrGroups, 4, false>(SCLGPU_SR_ARGS); break;
  }
#undef SCLGPU_SR_ARGS
  return cudaGetLastError();
}

}  // namespace sclgpu
