// sclgpu.cu -- host side of libsclgpu.so: context, launch logic, the C ABI of
// include/sclgpu.h.  Pure CUDA runtime; no torch types cross this boundary.
//
// Host-side arithmetic is limited to the AES-128 key schedule (PRG::create /
// aes128LoadKey, src/scl/util/prg.cc:54-101) and the constant limb images of the
// tensor-core kernels (powers of the evaluation points / Lagrange rows times 2^(8a),
// a few hundred field multiplications per (field, t, n), cached); everything that
// scales with the batch, and the Lagrange bases themselves, runs on the device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <new>
#include <set>
#include <string>
#include <vector>

#include "../../include/sclgpu.h"
#include "kernels.cuh"
#include "recover_c_big.cuh"
#include "recover_c_syndrome.cuh"
#include "matmul_tc.h"
#include "share_tc.h"
#include "host_stage.h"

using namespace sclgpu;

// ------------------------------------------------------------------ context
struct sclgpu_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;      // user stream for _dev entry points
  cudaStream_t pipe[2] = {nullptr, nullptr};  // host-pointer pipelines
  cudaEvent_t pipe_ev[2] = {nullptr, nullptr};
  uint32_t* d_t0 = nullptr;           // AES T0 table (256 words)
  int* d_flag = nullptr;              // lagrange "zero denominator" flag
  unsigned long long* d_count = nullptr;  // recover_d error counter
  void* d_partial = nullptr;          // dot/sum partials (kMaxPartials elements of 16 B)
  uint64_t launches = 0;
  std::string last_error;
  std::set<const void*> smem_opted;   // kernels with the 192 KiB opt-in done
  std::map<std::string, void*> basis_cache;  // (field, nodes, xs) -> device matrix
  std::map<uint32_t, void*> tc_bmat_cache;   // (t, n) -> Vandermonde limb image of k_share61_tc
  std::map<const void*, void*> rd_bmat_cache;  // Lagrange check matrix (device) -> its limb image for k_recover_d_tc
  std::map<const void*, RecBasis61> rec_basis_cache;  // Lagrange basis (device) -> 21-bit limbs for k_share_recover61
  bool tc_prepared = false;
  bool sr_prepared = false;
  cudaMemPool_t mempool = nullptr;    // private stream-ordered pool of the _dev entry points' scratch
  // device scratch of the host-pointer entry points (chunk buffers), kept across calls: a cudaMalloc /
  // cudaFree pair per buffer per call costs milliseconds and a device synchronisation
  std::vector<std::pair<void*, size_t>> pool;
  size_t pool_next = 0;
  // pageable host buffers (std::vector-backed SCL containers) go through a pinned ring + copy threads
  sclgpu::HostStager stager;
};

static constexpr int kMaxPartials = 2048;
static constexpr uint64_t kHostChunk = 1ull << 21;  // secrets per pipeline chunk

static int fail(sclgpu_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->last_error = msg;
  return code;
}
static int cuda_fail(sclgpu_ctx* ctx, cudaError_t e, const char* what) {
  return fail(ctx, e == cudaErrorMemoryAllocation ? SCLGPU_ENOMEM : SCLGPU_ECUDA,
              std::string(what) + ": " + cudaGetErrorString(e));
}
// No C++ exception crosses the C ABI: allocation failures of the host-side containers and anything else
// thrown below an entry point come back as an error code with a message.
template <class Fn>
static int guarded(sclgpu_ctx* ctx, Fn&& fn) {
  try {
    return fn();
  } catch (const std::bad_alloc&) {
    return fail(ctx, SCLGPU_ENOMEM, "host allocation failed");
  } catch (const std::exception& e) {
    return fail(ctx, SCLGPU_ECUDA, std::string("internal error: ") + e.what());
  } catch (...) {
    return fail(ctx, SCLGPU_ECUDA, "internal error");
  }
}

#define CK(call)                                                   \
  do {                                                             \
    cudaError_t e__ = (call);                                      \
    if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);     \
  } while (0)
#define CKL()                                                      \
  do {                                                             \
    ctx->launches++;                                               \
    cudaError_t e__ = cudaGetLastError();                          \
    if (e__ != cudaSuccess) return cuda_fail(ctx, e__, "launch");  \
  } while (0)
#define RET(x)                       \
  do {                               \
    int rc__ = (x);                  \
    if (rc__ != SCLGPU_OK) return rc__; \
  } while (0)

// ------------------------------------------------- AES-128 host key schedule
// FIPS-197 5.2 (== aes128LoadKey, prg.cc:54-75).  S-box from its definition.
static uint8_t g_sbox[256];
static uint32_t g_t0[256];
static std::once_flag g_aes_once;

static void aes_host_init_once() {
  uint8_t p = 1, q = 1;
  do {
    p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1b : 0));
    q ^= (uint8_t)(q << 1);
    q ^= (uint8_t)(q << 2);
    q ^= (uint8_t)(q << 4);
    if (q & 0x80) q ^= 0x09;
    const uint8_t x = (uint8_t)(q ^ (uint8_t)(q << 1 | q >> 7) ^ (uint8_t)(q << 2 | q >> 6) ^
                                (uint8_t)(q << 3 | q >> 5) ^ (uint8_t)(q << 4 | q >> 4));
    g_sbox[p] = (uint8_t)(x ^ 0x63);
  } while (p != 1);
  g_sbox[0] = 0x63;
  for (int i = 0; i < 256; ++i) {
    const uint8_t s = g_sbox[i];
    const uint8_t s2 = (uint8_t)((s << 1) ^ ((s >> 7) * 0x1b));
    const uint8_t s3 = (uint8_t)(s2 ^ s);
    g_t0[i] = (uint32_t)s2 | ((uint32_t)s << 8) | ((uint32_t)s << 16) | ((uint32_t)s3 << 24);
  }
}
static void aes_host_init() { std::call_once(g_aes_once, aes_host_init_once); }  // contexts may be created on several threads

static AesKey aes_expand(const uint8_t seed[16]) {
  static const uint8_t rcon[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1b, 0x36};
  AesKey k;
  for (int i = 0; i < 4; ++i) std::memcpy(&k.rk[i], seed + 4 * i, 4);
  for (int i = 4; i < 44; ++i) {
    uint32_t t = k.rk[i - 1];
    if ((i & 3) == 0) {
      t = (t >> 8) | (t << 24);
      t = (uint32_t)g_sbox[t & 0xff] | ((uint32_t)g_sbox[(t >> 8) & 0xff] << 8) |
          ((uint32_t)g_sbox[(t >> 16) & 0xff] << 16) | ((uint32_t)g_sbox[t >> 24] << 24);
      t ^= rcon[i / 4 - 1];
    }
    k.rk[i] = k.rk[i - 4] ^ t;
  }
  k.k8 = 1u << 8;
  k.k16 = 1u << 16;
  k.k24 = 1u << 24;
  return k;
}

// ------------------------------------------------------------ environment knobs
// Measurement / test switches.  Every variable is read ONCE per process, at its first use, and the
// answer is kept: no getenv on the dispatch paths.
static bool env_flag(const char* name) {
  static std::mutex mu;
  static std::map<std::string, bool> seen;
  std::lock_guard<std::mutex> lock(mu);
  auto it = seen.find(name);
  if (it != seen.end()) return it->second;
  const bool v = getenv(name) != nullptr;
  seen.emplace(name, v);
  return v;
}

static int env_int(const char* name, int dflt) {
  static std::mutex mu;
  static std::map<std::string, int> seen;
  std::lock_guard<std::mutex> lock(mu);
  auto it = seen.find(name);
  if (it != seen.end()) return it->second;
  const char* e = getenv(name);
  const int v = e ? atoi(e) : dflt;
  seen.emplace(name, v);
  return v;
}

// ------------------------------------------------------------ launch helpers
static int grid_for(const sclgpu_ctx* ctx, uint64_t work, int threads, int ctas_per_sm) {
  uint64_t g = (work + threads - 1) / threads;
  const uint64_t cap = (uint64_t)ctx->sm_count * ctas_per_sm;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

template <class K>
static int aes_opt_in(sclgpu_ctx* ctx, K kernel) {
  const void* f = reinterpret_cast<const void*>(kernel);
  if (ctx->smem_opted.count(f)) return SCLGPU_OK;
  CK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAesDynSmem));
  ctx->smem_opted.insert(f);
  return SCLGPU_OK;
}

// ------------------------------------------------------------ context / memory
extern "C" int sclgpu_init(int device, sclgpu_ctx** out) {
  if (!out) return SCLGPU_EINVAL;
  *out = nullptr;
  sclgpu_ctx* ctx = nullptr;
  try {
    ctx = new sclgpu_ctx();
  } catch (...) {
    return SCLGPU_ENOMEM;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
    delete ctx;
    return SCLGPU_ECUDA;  // no CPU fallback
  }
  ctx->device = device;
  cudaDeviceProp prop;
  if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
    delete ctx;
    return SCLGPU_ECUDA;
  }
  ctx->sm_count = prop.multiProcessorCount;
  ctx->cc_major = prop.major;
  ctx->cc_minor = prop.minor;
  aes_host_init();
  {  // stream-ordered scratch of the _dev entry points: a PRIVATE pool that keeps its memory between calls (the
     // device's default pool -- shared with every other cudaMallocAsync user of the process -- is left alone)
    cudaMemPoolProps props;
    std::memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&ctx->mempool, &props) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(ctx->mempool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      ctx->mempool = nullptr;
      cudaGetLastError();
    }
  }
  bool ok = cudaMalloc(&ctx->d_t0, sizeof(g_t0)) == cudaSuccess &&
            cudaMemcpy(ctx->d_t0, g_t0, sizeof(g_t0), cudaMemcpyHostToDevice) == cudaSuccess &&
            cudaMalloc(&ctx->d_flag, sizeof(int)) == cudaSuccess &&
            cudaMalloc(&ctx->d_count, sizeof(unsigned long long)) == cudaSuccess &&
            cudaMalloc(&ctx->d_partial, kMaxPartials * 16) == cudaSuccess;
  for (int i = 0; i < 2 && ok; ++i) {
    ok = cudaStreamCreateWithFlags(&ctx->pipe[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->pipe_ev[i], cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok) {
    sclgpu_destroy(ctx);
    return SCLGPU_ECUDA;
  }
  *out = ctx;
  return SCLGPU_OK;
}

// multi.cu: completion / teardown of this context's asynchronous call, if any
int sclgpu_async_complete(sclgpu_ctx* ctx, const char** error);
void sclgpu_async_release(sclgpu_ctx* ctx);

extern "C" void sclgpu_destroy(sclgpu_ctx* ctx) {
  if (!ctx) return;
  sclgpu_async_release(ctx);
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (auto& kv : ctx->basis_cache) cudaFree(kv.second);
  for (auto& kv : ctx->tc_bmat_cache) cudaFree(kv.second);
  for (auto& kv : ctx->rd_bmat_cache) cudaFree(kv.second);
  // the chunk buffers held secrets and shares of the caller's batches: wipe them before the memory goes back to the driver
  for (auto& pb : ctx->pool)
    if (pb.first) cudaMemset(pb.first, 0, pb.second);
  for (auto& pb : ctx->pool) cudaFree(pb.first);
  for (int i = 0; i < 2; ++i) {
    if (ctx->pipe[i]) cudaStreamDestroy(ctx->pipe[i]);
    if (ctx->pipe_ev[i]) cudaEventDestroy(ctx->pipe_ev[i]);
  }
  if (ctx->mempool) cudaMemPoolDestroy(ctx->mempool);
  cudaFree(ctx->d_t0);
  cudaFree(ctx->d_flag);
  cudaFree(ctx->d_count);
  cudaFree(ctx->d_partial);
  delete ctx;
}

extern "C" int sclgpu_set_stream(sclgpu_ctx* ctx, void* s) {
  if (!ctx) return SCLGPU_EINVAL;
  ctx->stream = (cudaStream_t)s;
  return SCLGPU_OK;
}
extern "C" int sclgpu_sync(sclgpu_ctx* ctx) {
  if (!ctx) return SCLGPU_EINVAL;
  const char* aerr = nullptr;
  const int arc = sclgpu_async_complete(ctx, &aerr);  // a pending *_async call finishes here
  CK(cudaSetDevice(ctx->device));
  CK(cudaStreamSynchronize(ctx->stream));
  if (arc != SCLGPU_OK) return fail(ctx, arc, aerr ? aerr : "asynchronous call failed");
  return SCLGPU_OK;
}
extern "C" int sclgpu_device_index(const sclgpu_ctx* ctx, int* device) {
  if (!ctx || !device) return SCLGPU_EINVAL;
  *device = ctx->device;
  return SCLGPU_OK;
}
extern "C" void sclgpu_set_error(sclgpu_ctx* ctx, const char* msg) {
  if (ctx) ctx->last_error = msg ? msg : "";
}
extern "C" const char* sclgpu_last_error(const sclgpu_ctx* ctx) {
  return ctx ? ctx->last_error.c_str() : "no context";
}
extern "C" const char* sclgpu_strerror(int code) {
  switch (code) {
    case SCLGPU_OK: return "ok";
    case SCLGPU_EINVAL: return "invalid argument";
    case SCLGPU_ELOGIC: return "logic error";
    case SCLGPU_EDETECT: return "error detected during recovery";
    case SCLGPU_ECUDA: return "CUDA failure or no usable device";
    case SCLGPU_ENOMEM: return "out of memory";
    case SCLGPU_ECORRECT: return "could not correct shares";
    default: return "unknown";
  }
}
extern "C" uint64_t sclgpu_launch_count(const sclgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
extern "C" int sclgpu_device_info(const sclgpu_ctx* ctx_, int* sm, int* maj, int* min, size_t* fr,
                                  size_t* tot) {
  sclgpu_ctx* ctx = const_cast<sclgpu_ctx*>(ctx_);
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  size_t f = 0, t = 0;
  CK(cudaMemGetInfo(&f, &t));
  if (sm) *sm = ctx->sm_count;
  if (maj) *maj = ctx->cc_major;
  if (min) *min = ctx->cc_minor;
  if (fr) *fr = f;
  if (tot) *tot = t;
  return SCLGPU_OK;
}
extern "C" int sclgpu_malloc(sclgpu_ctx* ctx, size_t bytes, void** p) {
  if (!ctx || !p) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMalloc(p, bytes ? bytes : 1));
  return SCLGPU_OK;
}
extern "C" int sclgpu_free(sclgpu_ctx* ctx, void* p) {
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaFree(p));
  return SCLGPU_OK;
}
extern "C" int sclgpu_host_alloc(sclgpu_ctx* ctx, size_t bytes, void** p) {
  if (!ctx || !p) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
  return SCLGPU_OK;
}
extern "C" int sclgpu_host_free(sclgpu_ctx* ctx, void* p) {
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaFreeHost(p));
  return SCLGPU_OK;
}
extern "C" int sclgpu_memcpy_h2d(sclgpu_ctx* ctx, void* d, const void* h, size_t bytes) {
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SCLGPU_OK;
}
extern "C" int sclgpu_memcpy_d2h(sclgpu_ctx* ctx, void* h, const void* d, size_t bytes) {
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return SCLGPU_OK;
}

extern "C" int sclgpu_memcpy_d2d(sclgpu_ctx* ctx, void* dst, const void* src, size_t bytes) {
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));  // also peer memory (UVA)
  return SCLGPU_OK;
}

// RAII device scratch
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() {
    if (p) cudaFree(p);
  }
  cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
  template <class T>
  T* as() { return reinterpret_cast<T*>(p); }
};

// Stream-ordered scratch of the device-pointer entry points: allocated and freed in stream order from the
// device's default pool, whose release threshold sclgpu_init raises, so that repeated calls reuse the same memory
// and no call pays a cudaMalloc / cudaFree pair (milliseconds, and a device-wide synchronisation).
struct StreamBuf {
  void* p = nullptr;
  sclgpu_ctx* ctx;
  cudaStream_t st;
  StreamBuf(sclgpu_ctx* c, cudaStream_t s) : ctx(c), st(s) {}
  StreamBuf(const StreamBuf&) = delete;
  StreamBuf& operator=(const StreamBuf&) = delete;
  ~StreamBuf() {
    if (p) cudaFreeAsync(p, st);
  }
  cudaError_t alloc(size_t bytes) {
    if (ctx && ctx->mempool) return cudaMallocFromPoolAsync(&p, bytes ? bytes : 1, ctx->mempool, st);
    return cudaMallocAsync(&p, bytes ? bytes : 1, st);
  }
  template <class T>
  T* as() { return reinterpret_cast<T*>(p); }
};

// Scratch from the context's pool: same interface as DevBuf, nothing freed on return.  A host entry
// point opens a PoolScope (calls on one context are serialised and end synchronised, so buffers handed
// out in one call are free again in the next).
static thread_local sclgpu_ctx* g_pool_ctx = nullptr;
struct PoolScope {
  explicit PoolScope(sclgpu_ctx* ctx) {
    g_pool_ctx = ctx;
    if (ctx) ctx->pool_next = 0;
  }
  ~PoolScope() {
    if (g_pool_ctx) {
      // error paths: nothing enqueued by the call -- kernels, async copies into the caller's buffers, staged copies --
      // outlives it (on the success path the streams are already idle and this costs nothing)
      cudaStreamSynchronize(g_pool_ctx->pipe[0]);
      cudaStreamSynchronize(g_pool_ctx->pipe[1]);
      g_pool_ctx->stager.drain();
    }
    g_pool_ctx = nullptr;
  }
};
struct PoolBuf {
  void* p = nullptr;
  cudaError_t alloc(size_t bytes) {
    sclgpu_ctx* ctx = g_pool_ctx;
    if (!ctx) return cudaErrorInvalidValue;
    if (bytes == 0) bytes = 1;
    if (ctx->pool_next == ctx->pool.size()) ctx->pool.emplace_back(nullptr, 0);
    auto& slot = ctx->pool[ctx->pool_next++];
    if (slot.second < bytes) {
      if (slot.first) cudaFree(slot.first);
      slot = {nullptr, 0};
      void* q = nullptr;
      cudaError_t e = cudaMalloc(&q, bytes);
      if (e != cudaSuccess) return e;
      slot = {q, bytes};
    }
    p = slot.first;
    return cudaSuccess;
  }
  template <class T>
  T* as() { return reinterpret_cast<T*>(p); }
};

// ============================================================ device-level ops
// (stream passed explicitly so the host pipelines can use their own)

static int prg_bytes_on(sclgpu_ctx* ctx, cudaStream_t st, const uint8_t seed[16], uint64_t first_block,
                        uint64_t n_bytes, uint8_t* d_out) {
  if (n_bytes == 0) return SCLGPU_OK;
  RET(aes_opt_in(ctx, k_prg_bytes));
  const AesKey key = aes_expand(seed);
  const int grid = grid_for(ctx, (n_bytes + 15) / 16, kAesThreads, 1);
  k_prg_bytes<<<grid, kAesThreads, kAesDynSmem, st>>>(key, ctx->d_t0, first_block, n_bytes, d_out);
  CKL();
  return SCLGPU_OK;
}

template <class F, bool ONE>
static int random_on(sclgpu_ctx* ctx, cudaStream_t st, const uint8_t seed[16], uint64_t first_block,
                     uint64_t n, typename F::E* d_out) {
  if (n == 0) return SCLGPU_OK;
  RET(aes_opt_in(ctx, k_random<F, ONE>));
  const AesKey key = aes_expand(seed);
  const int grid = grid_for(ctx, n, kAesThreads, 1);
  k_random<F, ONE><<<grid, kAesThreads, kAesDynSmem, st>>>(key, ctx->d_t0, first_block, n, d_out);
  CKL();
  return SCLGPU_OK;
}

template <class F>
static int from_bytes_on(sclgpu_ctx* ctx, cudaStream_t st, const uint8_t* d_bytes, uint64_t n,
                         typename F::E* d_out) {
  if (n == 0) return SCLGPU_OK;
  k_from_bytes<F><<<grid_for(ctx, n, 256, 8), 256, 0, st>>>(d_bytes, n, d_out);
  CKL();
  return SCLGPU_OK;
}

template <class F, int T>
static int share_fused_launch(sclgpu_ctx* ctx, cudaStream_t st, const AesKey& key, uint64_t first_block,
                              const typename F::E* d_secrets, uint64_t N, uint32_t n,
                              typename F::E* d_out, uint64_t si, uint64_t sj) {
  RET(aes_opt_in(ctx, k_share_fused<F, T>));
  const int grid = grid_for(ctx, N, kAesThreads, 1);
  k_share_fused<F, T><<<grid, kAesThreads, kAesDynSmem, st>>>(key, ctx->d_t0, first_block, d_secrets, N,
                                                             n, d_out, si, sj);
  CKL();
  return SCLGPU_OK;
}

// tensor-core Fp61 kernel (share_tc.cu): t <= 15, n <= 32.  The B operand holds the
// bytes of C[i][k][a] = (i+1)^k * 2^(8a) mod p (the Vandermonde entry of party i,
// matrix.h:445-460, pre-multiplied by the weight of coefficient byte a): small
// constants, computed here once per (t, n) and kept on the device.
static int g_share_tc = -1;
static bool share_tc_enabled() {
  if (g_share_tc < 0) {
    // 0 = integer-pipe kernel (k_share61); tcgen05 kernels: 1 = A operand in shared memory (3 groups),
    // 2 / 3 = A operand in tensor memory with 4 / 5 groups of warps (3 is the default),
    // 4 = warp-specialised: 4 producer groups (AES) + 2 consumer groups (MMA, epilogue)
    const char* e = getenv("SCLGPU_SHARE_TC");
    g_share_tc = e ? atoi(e) : 3;
    if (g_share_tc < 0 || g_share_tc > 4) g_share_tc = 3;
  }
  return g_share_tc != 0;
}

// B image for field F: row r = party*BYTES + limb, column kk = coeff*BYTES + byte
template <class F>
static int share_tc_bmat(sclgpu_ctx* ctx, cudaStream_t st, uint32_t t, uint32_t n, const void** d_bmat) {
  typedef typename F::E E;
  constexpr uint32_t EB = F::BYTES;
  const uint32_t key = ((uint32_t)EB << 16) | (t << 8) | n;
  auto it = ctx->tc_bmat_cache.find(key);
  if (it != ctx->tc_bmat_cache.end()) {
    *d_bmat = it->second;
    return SCLGPU_OK;
  }
  std::vector<uint8_t> img(kTcBmatBytes, 0);
  for (uint32_t i = 0; i < n; ++i) {
    E pw = F::one();  // (i+1)^k
    for (uint32_t k = 0; k <= t; ++k) {
      E c = pw;       // (i+1)^k * 2^(8a)
      for (uint32_t a = 0; a < EB; ++a) {
        uint8_t bytes[16];
        std::memcpy(bytes, &c, EB);  // little-endian canonical residue = its 8-bit limbs
        for (uint32_t s = 0; s < EB; ++s) img[tc_bmat_offset(i * EB + s, k * EB + a)] = bytes[s];
        c = F::mul(c, F::from_u32(256));
      }
      pw = F::mul(pw, F::from_u32(i + 1));
    }
  }
  void* d = nullptr;
  CK(cudaMalloc(&d, kTcBmatBytes));
  cudaError_t e = cudaMemcpyAsync(d, img.data(), kTcBmatBytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // img is a local
  if (e != cudaSuccess) {
    cudaFree(d);
    return cuda_fail(ctx, e, "share_tc constants");
  }
  ctx->tc_bmat_cache[key] = d;
  *d_bmat = d;
  return SCLGPU_OK;
}

template <class F>
static int share_tc_on(sclgpu_ctx* ctx, cudaStream_t st, const AesKey& key, uint64_t first_block,
                       const typename F::E* d_secrets, uint64_t N, uint32_t t, uint32_t n, typename F::E* d_out,
                       uint64_t si, uint64_t sj) {
  if (!ctx->tc_prepared) {
    CK(share_tc_prepare());
    ctx->tc_prepared = true;
  }
  const void* d_bmat = nullptr;
  RET(share_tc_bmat<F>(ctx, st, t, n, &d_bmat));
  const uint64_t tiles = (N + 127) / 128;
  int variant = g_share_tc;
  if (F::BYTES == 16 && variant == 1) variant = 3;  // the shared-memory-A kernel exists for Fp61 only
  const int groups = tc_variant_groups(variant);
  const int share_sms = std::min(ctx->sm_count, std::max(1, env_int("SCLGPU_SHARE_SMS", ctx->sm_count)));
  const int grid = (int)std::min<uint64_t>((tiles + groups - 1) / groups, (uint64_t)share_sms);
  ctx->launches++;
  cudaError_t e;
  if constexpr (F::BYTES == 8) {
    e = share61_tc_launch(variant, st, grid, key, ctx->d_t0, d_bmat, first_block, d_secrets, N, t, n, d_out, si, sj);
  } else {
    e = share127_tc_launch(variant, st, grid, key, ctx->d_t0, d_bmat, first_block, d_secrets, N, t, n, d_out, si, sj);
  }
  if (e != cudaSuccess) return cuda_fail(ctx, e, "launch");
  return SCLGPU_OK;
}

// tuned Fp61 kernel (k_share61): t <= 15, n <= 65535
static int g_addmode = -1;
static int share61_addmode() {
  if (g_addmode < 0) {
    const char* e = getenv("SCLGPU_ADDMODE");  // tuning knob: vl by 0 shift (ALU), 1 mul.hi (FMA), 2 alternate
    g_addmode = e ? atoi(e) : 0;
    if (g_addmode < 0 || g_addmode > 2) g_addmode = 0;
  }
  return g_addmode;
}

template <int T, int ADDMODE>
static int share61_launch(sclgpu_ctx* ctx, cudaStream_t st, const AesKey& key, uint64_t first_block,
                          const uint64_t* d_secrets, uint64_t N, uint32_t n, uint64_t* d_out, uint64_t si,
                          uint64_t sj) {
  RET(aes_opt_in(ctx, k_share61<T, ADDMODE>));
  const int grid = grid_for(ctx, N, kAesThreads, 1);
  k_share61<T, ADDMODE><<<grid, kAesThreads, kAesDynSmem, st>>>(key, ctx->d_t0, first_block, d_secrets, N, n,
                                                              d_out, si, sj, 1u << 29);
  CKL();
  return SCLGPU_OK;
}

template <int T>
static int share61_mode(sclgpu_ctx* ctx, cudaStream_t st, const AesKey& key, uint64_t first_block,
                        const uint64_t* d_secrets, uint64_t N, uint32_t n, uint64_t* d_out, uint64_t si,
                        uint64_t sj) {
  switch (share61_addmode()) {
    case 0: return share61_launch<T, 0>(ctx, st, key, first_block, d_secrets, N, n, d_out, si, sj);
    case 1: return share61_launch<T, 1>(ctx, st, key, first_block, d_secrets, N, n, d_out, si, sj);
    default: return share61_launch<T, 2>(ctx, st, key, first_block, d_secrets, N, n, d_out, si, sj);
  }
}

template <class F>
static constexpr int max_fused_t() {
  return F::BYTES == 8 ? 16 : 8;
}

template <class F>
static int share_coeffs_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* d_coeffs, uint64_t N,
                           uint32_t t, uint32_t n, typename F::E* d_out, uint64_t si, uint64_t sj) {
  if (N == 0 || n == 0) return SCLGPU_OK;
  const bool fits = F::BYTES == 8 ? (t <= kTcMaxT && n <= kTcMaxParties) : (t <= kTcMaxT127 && n <= kTcMaxParties127);
  if (fits && share_tc_enabled() && !env_flag("SCLGPU_SHARE_GENERIC")) {
    const void* d_bmat = nullptr;
    RET(share_tc_bmat<F>(ctx, st, t, n, &d_bmat));
    ctx->launches++;
    cudaError_t e;
    if constexpr (F::BYTES == 8) {
      e = share61_coeffs_tc_launch(st, ctx->sm_count, d_bmat, d_coeffs, N, t, n, d_out, si, sj);
    } else {
      e = share127_coeffs_tc_launch(st, ctx->sm_count, d_bmat, d_coeffs, N, t, n, d_out, si, sj);
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch");
    return SCLGPU_OK;
  }
  k_share_coeffs<F><<<grid_for(ctx, N, 256, 8), 256, 0, st>>>(d_coeffs, N, t, n, d_out, si, sj);
  CKL();
  return SCLGPU_OK;
}

// party-major / strided share of N secrets (out index = i*si + j*sj)
template <class F>
static int share_strided_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* d_secrets, uint64_t N,
                            uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t first_block,
                            typename F::E* d_out, uint64_t si, uint64_t sj) {
  typedef typename F::E E;
  if (N == 0 || n == 0) return SCLGPU_OK;
  const AesKey key = aes_expand(seed);
  if constexpr (F::BYTES == 8) {
    if (t <= kTcMaxT && n <= kTcMaxParties && share_tc_enabled() && !env_flag("SCLGPU_SHARE_GENERIC"))
      return share_tc_on<F61>(ctx, st, key, first_block, d_secrets, N, t, n, d_out, si, sj);
    if (t <= 15 && n <= 0xFFFFu && !env_flag("SCLGPU_SHARE_GENERIC")) {
#define SCLGPU_CASE61(TT) \
  case TT: return share61_mode<TT>(ctx, st, key, first_block, d_secrets, N, n, d_out, si, sj);
      switch (t) {
        SCLGPU_CASE61(0) SCLGPU_CASE61(1) SCLGPU_CASE61(2) SCLGPU_CASE61(3) SCLGPU_CASE61(4) SCLGPU_CASE61(5)
        SCLGPU_CASE61(6) SCLGPU_CASE61(7) SCLGPU_CASE61(8) SCLGPU_CASE61(9) SCLGPU_CASE61(10) SCLGPU_CASE61(11)
        SCLGPU_CASE61(12) SCLGPU_CASE61(13) SCLGPU_CASE61(14) SCLGPU_CASE61(15)
        default: break;
      }
#undef SCLGPU_CASE61
    }
  }
  if constexpr (F::BYTES == 16) {
    if (t <= kTcMaxT127 && n <= kTcMaxParties127 && share_tc_enabled() && !env_flag("SCLGPU_SHARE_GENERIC"))
      return share_tc_on<F127>(ctx, st, key, first_block, d_secrets, N, t, n, d_out, si, sj);
  }
#define SCLGPU_CASE(TT) \
  case TT: return share_fused_launch<F, (TT <= max_fused_t<F>() ? TT : 0)>(ctx, st, key, first_block, d_secrets, N, n, d_out, si, sj);
  if ((int)t <= max_fused_t<F>()) {
    switch (t) {
      SCLGPU_CASE(0) SCLGPU_CASE(1) SCLGPU_CASE(2) SCLGPU_CASE(3) SCLGPU_CASE(4) SCLGPU_CASE(5)
      SCLGPU_CASE(6) SCLGPU_CASE(7) SCLGPU_CASE(8) SCLGPU_CASE(9) SCLGPU_CASE(10) SCLGPU_CASE(11)
      SCLGPU_CASE(12) SCLGPU_CASE(13) SCLGPU_CASE(14) SCLGPU_CASE(15) SCLGPU_CASE(16)
      default: break;
    }
  }
#undef SCLGPU_CASE
  // any other threshold: keystream -> coefficient planes -> evaluation, in
  // chunks that keep the planes below ~1 GiB
  const uint64_t B = ((uint64_t)(t + 1) * F::BYTES + 15) / 16;
  uint64_t chunk = (1ull << 30) / ((uint64_t)(t + 1) * sizeof(E));
  if (chunk < 1024) chunk = 1024;
  if (chunk > N) chunk = N;
  StreamBuf planes(ctx, st);
  CK(planes.alloc(chunk * (uint64_t)(t + 1) * sizeof(E)));
  RET(aes_opt_in(ctx, k_expand_coeffs<F>));
  for (uint64_t c0 = 0; c0 < N; c0 += chunk) {
    const uint64_t nc = std::min(chunk, N - c0);
    k_expand_coeffs<F><<<grid_for(ctx, nc, kAesThreads, 1), kAesThreads, kAesDynSmem, st>>>(
        key, ctx->d_t0, first_block + c0 * B, d_secrets + c0, nc, t, 1u, planes.as<E>());
    CKL();
    RET(share_coeffs_on<F>(ctx, st, planes.as<E>(), nc, t, n, d_out + c0 * sj, si, sj));
  }
  return SCLGPU_OK;  // planes goes back to the pool in stream order
}

template <class E>
static int transpose_on(sclgpu_ctx* ctx, cudaStream_t st, const E* d_in, uint64_t rows, uint64_t cols,
                        E* d_out) {
  if (rows == 0 || cols == 0) return SCLGPU_OK;
  const uint64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32);
  const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 16);
  k_transpose<E><<<grid, 256, 0, st>>>(d_in, rows, cols, d_out);
  CKL();
  return SCLGPU_OK;
}

// Lagrange rows on the device, cached per (field, nodes, xs).  nodes == nullptr
// -> 1..m.  Returns a device matrix rows x m.
template <class F>
static int basis_rows(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* nodes, uint32_t m,
                      const typename F::E* xs, uint32_t rows, const typename F::E** d_mat) {
  typedef typename F::E E;
  std::vector<E> hn(m), hx(rows);
  for (uint32_t i = 0; i < m; ++i) hn[i] = nodes ? nodes[i] : F::from_u32(i + 1);
  for (uint32_t r = 0; r < rows; ++r) hx[r] = xs[r];
  std::string key(1, (char)F::BYTES);
  key.append(reinterpret_cast<const char*>(&m), 4);
  key.append(reinterpret_cast<const char*>(hn.data()), (size_t)m * sizeof(E));
  key.append(reinterpret_cast<const char*>(hx.data()), (size_t)rows * sizeof(E));
  auto it = ctx->basis_cache.find(key);
  if (it != ctx->basis_cache.end()) {
    *d_mat = reinterpret_cast<const E*>(it->second);
    return SCLGPU_OK;
  }
  DevBuf dn, dx;
  void* dm = nullptr;
  CK(dn.alloc((size_t)m * sizeof(E)));
  CK(dx.alloc((size_t)rows * sizeof(E)));
  CK(cudaMalloc(&dm, std::max<size_t>((size_t)rows * m * sizeof(E), 16)));
  cudaError_t e = cudaMemcpyAsync(dn.p, hn.data(), (size_t)m * sizeof(E), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(dx.p, hx.data(), (size_t)rows * sizeof(E), cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), st);
  if (e != cudaSuccess) {
    cudaFree(dm);
    return cuda_fail(ctx, e, "basis upload");
  }
  if (m > 0 && rows > 0) {
    k_lagrange_rows<F><<<rows, (int)std::min<uint32_t>(std::max<uint32_t>(m, 32), 256), 0, st>>>(
        dn.as<E>(), m, dx.as<E>(), reinterpret_cast<E*>(dm), ctx->d_flag);
    ctx->launches++;
  }
  int bad = 0;
  e = cudaMemcpyAsync(&bad, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) {
    cudaFree(dm);
    return cuda_fail(ctx, e, "lagrange");
  }
  if (bad) {
    cudaFree(dm);
    return fail(ctx, SCLGPU_ELOGIC, "0 not invertible modulo prime");
  }
  if (ctx->basis_cache.size() > 64) {
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->basis_cache) cudaFree(kv.second);
    ctx->basis_cache.clear();
    for (auto& kv : ctx->rd_bmat_cache) cudaFree(kv.second);  // keyed by the pointers just freed
    ctx->rd_bmat_cache.clear();
    ctx->rec_basis_cache.clear();
  }
  ctx->basis_cache[key] = dm;
  *d_mat = reinterpret_cast<const E*>(dm);
  return SCLGPU_OK;
}

template <class F>
static int recover_d_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* d_shares, uint64_t N,
                        uint64_t si, uint64_t sj, uint32_t m, uint32_t n_checks,
                        const typename F::E* d_mat, typename F::E* d_out, uint8_t* d_err);

template <class F>
static int recover_p_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* d_shares, uint64_t N,
                        uint32_t n, uint64_t si, uint64_t sj, const typename F::E* d_basis,
                        typename F::E* d_out, const GatherDst* gather = nullptr) {
  if (N == 0) return SCLGPU_OK;
  GatherDst gd;
  std::memset(&gd, 0, sizeof(gd));
  if (gather) gd = *gather;
  if constexpr (F::BYTES == 8) {
    // party-major planes, the device-native layout: HBM-bound kernel
    if (env_flag("SCLGPU_RECOVER61_TC") && recover_d_tc_fits<F>(n, 0))
      return recover_d_on<F>(ctx, st, d_shares, N, si, sj, n, 0, d_basis, d_out, nullptr);
    if (sj == 1 && n >= 1 && n <= 2048 && !env_flag("SCLGPU_RECOVER_GENERIC")) {
      uintptr_t align = reinterpret_cast<uintptr_t>(d_shares) | reinterpret_cast<uintptr_t>(d_out);
      for (uint32_t g = 0; g < gd.count; ++g) align |= reinterpret_cast<uintptr_t>(gd.dst[g]);
      const bool vec2 = (N % 2 == 0) && (si % 2 == 0) && (align & 15) == 0;
      const size_t lsm = (size_t)n * 16;
      if (vec2) {
        k_recover61_pm<2><<<std::min(grid_for(ctx, N / 2, 256, 3), 3 * std::max(1, env_int("SCLGPU_RECOVER_SMS", ctx->sm_count))), 256, lsm, st>>>(d_shares, N, n, si, d_basis, d_out, gd);
      } else {
        k_recover61_pm<1><<<grid_for(ctx, N, 256, 4), 256, lsm, st>>>(d_shares, N, n, si, d_basis, d_out, gd);
      }
      CKL();
      return SCLGPU_OK;
    }
  }
  if (gd.count) return fail(ctx, SCLGPU_EINVAL, "gathered reconstruction needs party-major Fp61 planes");
  if constexpr (F::BYTES == 16) {
    // Fp127: the inner product as a one-row limb product on the tensor cores (k_recover_d_tc without checks)
    if (recover_d_tc_fits<F>(n, 0) && !env_flag("SCLGPU_RECOVER_GENERIC"))
      return recover_d_on<F>(ctx, st, d_shares, N, si, sj, n, 0, d_basis, d_out, nullptr);
  }
  // basis staged in shared memory: up to 200 KiB (n <= 25600 for Fp61, 12800 for Fp127; documented in sclgpu.h)
  const size_t smem = (size_t)n * sizeof(typename F::E);
  if (smem > 200 * 1024) return fail(ctx, SCLGPU_EINVAL, "recover_p: more than 200 KiB of Lagrange basis");
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_recover_p<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_recover_p<F><<<grid_for(ctx, N, 256, 8), 256, smem, st>>>(d_shares, N, n, si, sj, d_basis, d_out);
  CKL();
  return SCLGPU_OK;
}

// the (alphas, t, d, x) -> check matrix part of shamirRecoverD (shamir.h:117-131, 137-139)
template <class F>
static int recover_d_matrix(sclgpu_ctx* ctx, cudaStream_t st, uint32_t n_given, uint32_t& t,
                            const typename F::E* alphas, uint32_t& n_alphas, uint32_t& d,
                            const typename F::E* x, uint32_t& m, uint32_t& n_checks,
                            const typename F::E** d_mat) {
  typedef typename F::E E;
  std::vector<E> al;
  E xx = F::zero();
  if (alphas == nullptr) {  // shamir.h:152-155
    n_alphas = 2 * t + 1;
    d = t;
    al.resize(n_alphas);
    for (uint32_t i = 0; i < n_alphas; ++i) al[i] = F::from_u32(i + 1);
  } else {
    al.assign(alphas, alphas + n_alphas);
    if (x) xx = *x;
  }
  if ((uint64_t)n_given < (uint64_t)d + t || (uint64_t)n_alphas < (uint64_t)d + t)
    return fail(ctx, SCLGPU_ELOGIC, "not enough shares provided to detect errors");
  // the interpolation reads shares and nodes 0..d: with t = 0 the reference's own check lets d + 1 > n_given through
  // and reads past the end of both vectors (shamir.h:125-127); here that is an error code
  if ((uint64_t)n_given < (uint64_t)d + 1 || (uint64_t)n_alphas < (uint64_t)d + 1)
    return fail(ctx, SCLGPU_ELOGIC, "not enough shares provided to detect errors");
  m = d + 1;
  n_checks = (d + t > m) ? d + t - m : 0;
  std::vector<E> xs(n_checks + 1);
  for (uint32_t r = 0; r < n_checks; ++r) xs[r] = al[m + r];
  xs[n_checks] = xx;
  return basis_rows<F>(ctx, st, al.data(), m, xs.data(), n_checks + 1, d_mat);
}

template <class F>
static int recover_d_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* d_shares, uint64_t N,
                        uint64_t si, uint64_t sj, uint32_t m, uint32_t n_checks,
                        const typename F::E* d_mat, typename F::E* d_out, uint8_t* d_err) {
  if (N == 0) return SCLGPU_OK;
  if (recover_d_tc_fits<F>(m, n_checks) && !env_flag("SCLGPU_RECOVER_GENERIC")) {
    // tensor-core kernel: limb image of the (n_checks+1) x m matrix, row r*BYTES+s, column k*BYTES+a
    typedef typename F::E E;
    constexpr uint32_t EB = F::BYTES;
    void* d_img = nullptr;
    auto it = ctx->rd_bmat_cache.find(d_mat);
    if (it != ctx->rd_bmat_cache.end()) {
      d_img = it->second;
    } else {
      const uint32_t rows = n_checks + 1;
      std::vector<E> hm((size_t)rows * m);
      CK(cudaMemcpyAsync(hm.data(), d_mat, hm.size() * sizeof(E), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      std::vector<uint8_t> img(kTcBmatBytes, 0);
      for (uint32_t r = 0; r < rows; ++r)
        for (uint32_t k = 0; k < m; ++k) {
          E c = hm[(size_t)r * m + k];
          for (uint32_t a = 0; a < EB; ++a) {
            uint8_t bytes[16];
            std::memcpy(bytes, &c, EB);
            for (uint32_t s = 0; s < EB; ++s) img[tc_rd_offset(r * EB + s, k * EB + a)] = bytes[s];
            c = F::mul(c, F::from_u32(256));
          }
        }
      CK(cudaMalloc(&d_img, kTcBmatBytes));
      cudaError_t e = cudaMemcpyAsync(d_img, img.data(), kTcBmatBytes, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) {
        cudaFree(d_img);
        return cuda_fail(ctx, e, "recover_d_tc constants");
      }
      ctx->rd_bmat_cache[d_mat] = d_img;
    }
    ctx->launches++;
    cudaError_t e;
    if constexpr (EB == 8) {
      e = recover_d61_tc_launch(st, std::min(ctx->sm_count, std::max(1, env_int("SCLGPU_RECOVER_SMS", ctx->sm_count))), d_img, d_shares, N, si, sj, m, n_checks, d_out, d_err, d_err ? ctx->d_count : nullptr);
    } else {
      e = recover_d127_tc_launch(st, ctx->sm_count, d_img, d_shares, N, si, sj, m, n_checks, d_out, d_err, d_err ? ctx->d_count : nullptr);
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch");
    return SCLGPU_OK;
  }
  const size_t smem = (size_t)(n_checks + 1) * m * sizeof(typename F::E);
  if (smem > 200 * 1024) return fail(ctx, SCLGPU_EINVAL, "recover_d: check matrix exceeds shared memory");
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_recover_d<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_recover_d<F><<<grid_for(ctx, N, 256, 8), 256, smem, st>>>(d_shares, N, si, sj, m, n_checks, d_mat,
                                                              d_out, d_err, ctx->d_count);
  CKL();
  return SCLGPU_OK;
}

static void strides_for(int layout, uint64_t N, uint32_t n, uint64_t& si, uint64_t& sj) {
  if (layout == SCLGPU_PARTY_MAJOR) {
    si = N;
    sj = 1;
  } else {
    si = 1;
    sj = n;
  }
}

// ================================================================ C ABI: PRG
static int sclgpu_prg_expand_dev_impl(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                                     uint64_t n_bytes, uint8_t* d_out) {
  if (!ctx || !seed || (!d_out && n_bytes)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return prg_bytes_on(ctx, ctx->stream, seed, first_block, n_bytes, d_out);
}

static int sclgpu_prg_expand_impl(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block,
                                 uint64_t n_bytes, uint8_t* out) {
  if (!ctx || !seed || (!out && n_bytes)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  if (n_bytes == 0) return SCLGPU_OK;
  const uint64_t chunk = 256ull << 20;  // bytes, multiple of 16
  PoolScope pool_scope(ctx);
  PoolBuf buf[2];
  CK(buf[0].alloc(std::min(chunk, n_bytes)));
  if (n_bytes > chunk) CK(buf[1].alloc(std::min(chunk, n_bytes - chunk)));
  int k = 0;
  for (uint64_t off = 0; off < n_bytes; off += chunk, k ^= 1) {
    const uint64_t nb = std::min(chunk, n_bytes - off);
    cudaStream_t st = ctx->pipe[k];
    RET(prg_bytes_on(ctx, st, seed, first_block + off / 16, nb, buf[k].as<uint8_t>()));
    CK(ctx->stager.d2h(st, out + off, buf[k].p, nb));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}

// ================================== generic host wrappers (upload / run / download)
// Small helper for the host entry points whose buffers comfortably fit on the
// device at once: inputs are copied up on pipe[0], the op runs there, outputs
// come back, one synchronise at the end.
struct HostOp {
  sclgpu_ctx* ctx;
  cudaStream_t st;
  std::vector<void*> bufs;
  explicit HostOp(sclgpu_ctx* c) : ctx(c), st(c->pipe[0]) {}
  ~HostOp() {
    cudaStreamSynchronize(st);
    ctx->stager.drain();
    for (void* p : bufs) cudaFree(p);
  }
  int up(const void* h, size_t bytes, void** d) {
    CK(cudaMalloc(d, bytes ? bytes : 1));
    bufs.push_back(*d);
    if (bytes) CK(ctx->stager.h2d(st, *d, h, bytes));
    return SCLGPU_OK;
  }
  int dev(size_t bytes, void** d) {
    CK(cudaMalloc(d, bytes ? bytes : 1));
    bufs.push_back(*d);
    return SCLGPU_OK;
  }
  int down(void* h, const void* d, size_t bytes) {
    if (bytes) CK(ctx->stager.d2h(st, h, d, bytes));
    CK(cudaStreamSynchronize(st));
    CK(ctx->stager.drain());
    return SCLGPU_OK;
  }
};

// --------------------------------------------------------------- random etc.
template <class F, bool ONE>
static int random_host(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n,
                       void* out) {
  typedef typename F::E E;
  if (!ctx || !seed || (!out && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return SCLGPU_OK;
  const uint64_t chunk = 1ull << 25;  // elements; even, so Fp61 chunks start on a block boundary
  const uint64_t per_block = (F::BYTES == 16 || ONE) ? 1 : 2;
  PoolScope pool_scope(ctx);
  PoolBuf buf[2];
  CK(buf[0].alloc(std::min(chunk, n) * sizeof(E)));
  if (n > chunk) CK(buf[1].alloc(std::min(chunk, n - chunk) * sizeof(E)));
  int k = 0;
  for (uint64_t off = 0; off < n; off += chunk, k ^= 1) {
    const uint64_t nc = std::min(chunk, n - off);
    cudaStream_t st = ctx->pipe[k];
    RET((random_on<F, ONE>(ctx, st, seed, first_block + off / per_block, nc, buf[k].as<E>())));
    CK(ctx->stager.d2h(st, reinterpret_cast<E*>(out) + off, buf[k].p, nc * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}

template <class F, bool ONE>
static int random_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n,
                      void* d_out) {
  if (!ctx || !seed || (!d_out && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return random_on<F, ONE>(ctx, ctx->stream, seed, first_block, n, reinterpret_cast<typename F::E*>(d_out));
}

template <class F>
static int from_bytes_host(sclgpu_ctx* ctx, const uint8_t* bytes, uint64_t n, void* out) {
  typedef typename F::E E;
  if (!ctx || ((!bytes || !out) && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return SCLGPU_OK;
  HostOp op(ctx);
  void *db, *dout;
  RET(op.up(bytes, n * F::BYTES, &db));
  RET(op.dev(n * sizeof(E), &dout));
  RET(from_bytes_on<F>(ctx, op.st, (const uint8_t*)db, n, (E*)dout));
  return op.down(out, dout, n * sizeof(E));
}

extern "C" int sclgpu_fp61_from_bytes(sclgpu_ctx* c, const uint8_t* b, uint64_t n, uint64_t* o) { return guarded(c, [&] { return from_bytes_host<F61>(c, b, n, o); }); }
extern "C" int sclgpu_fp127_from_bytes(sclgpu_ctx* c, const uint8_t* b, uint64_t n, void* o) { return guarded(c, [&] { return from_bytes_host<F127>(c, b, n, o); }); }
extern "C" int sclgpu_fp61_random(sclgpu_ctx* c, const uint8_t s[16], uint64_t fb, uint64_t n, uint64_t* o) { return guarded(c, [&] { return random_host<F61, false>(c, s, fb, n, o); }); }
extern "C" int sclgpu_fp127_random(sclgpu_ctx* c, const uint8_t s[16], uint64_t fb, uint64_t n, void* o) { return guarded(c, [&] { return random_host<F127, false>(c, s, fb, n, o); }); }
extern "C" int sclgpu_fp61_ff_random(sclgpu_ctx* c, const uint8_t s[16], uint64_t fb, uint64_t n, uint64_t* o) { return guarded(c, [&] { return random_host<F61, true>(c, s, fb, n, o); }); }
extern "C" int sclgpu_fp127_ff_random(sclgpu_ctx* c, const uint8_t s[16], uint64_t fb, uint64_t n, void* o) { return guarded(c, [&] { return random_host<F127, true>(c, s, fb, n, o); }); }
extern "C" int sclgpu_fp61_random_dev(sclgpu_ctx* c, const uint8_t s[16], uint64_t fb, uint64_t n, uint64_t* o) { return guarded(c, [&] { return random_dev<F61, false>(c, s, fb, n, o); }); }
extern "C" int sclgpu_fp127_random_dev(sclgpu_ctx* c, const uint8_t s[16], uint64_t fb, uint64_t n, void* o) { return guarded(c, [&] { return random_dev<F127, false>(c, s, fb, n, o); }); }
extern "C" int sclgpu_fp61_ff_random_dev(sclgpu_ctx* c, const uint8_t s[16], uint64_t fb, uint64_t n, uint64_t* o) { return guarded(c, [&] { return random_dev<F61, true>(c, s, fb, n, o); }); }
extern "C" int sclgpu_fp127_ff_random_dev(sclgpu_ctx* c, const uint8_t s[16], uint64_t fb, uint64_t n, void* o) { return guarded(c, [&] { return random_dev<F127, true>(c, s, fb, n, o); }); }
static int sclgpu_fp61_from_bytes_dev_impl(sclgpu_ctx* ctx, const uint8_t* b, uint64_t n, uint64_t* o) {
  if (!ctx || ((!b || !o) && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return from_bytes_on<F61>(ctx, ctx->stream, b, n, o);
}
static int sclgpu_fp127_from_bytes_dev_impl(sclgpu_ctx* ctx, const uint8_t* b, uint64_t n, void* o) {
  if (!ctx || ((!b || !o) && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return from_bytes_on<F127>(ctx, ctx->stream, b, n, (E127*)o);
}

// ------------------------------------------------------------------ share
template <class F>
static int share_dev(sclgpu_ctx* ctx, const void* d_secrets, uint64_t N, uint32_t t, uint32_t n,
                     const uint8_t seed[16], uint64_t first_block, void* d_shares, int layout) {
  typedef typename F::E E;
  if (!ctx || !seed || ((!d_secrets || !d_shares) && N && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n >= (1u << 31)) return fail(ctx, SCLGPU_EINVAL, "n too large");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  CK(cudaSetDevice(ctx->device));
  if (N == 0 || n == 0) return SCLGPU_OK;
  const E* sec = reinterpret_cast<const E*>(d_secrets);
  E* out = reinterpret_cast<E*>(d_shares);
  if (layout == SCLGPU_PARTY_MAJOR)
    return share_strided_on<F>(ctx, ctx->stream, sec, N, t, n, seed, first_block, out, N, 1);
  // secret-major: evaluate party-major into a chunk buffer, then transpose the
  // chunk into place (coalesced on both sides)
  const uint64_t B = ((uint64_t)(t + 1) * F::BYTES + 15) / 16;
  uint64_t chunk = std::max<uint64_t>((512ull << 20) / ((uint64_t)n * sizeof(E)), 1024);
  if (chunk > N) chunk = N;
  StreamBuf tmp(ctx, ctx->stream);
  CK(tmp.alloc(chunk * n * sizeof(E)));
  for (uint64_t c0 = 0; c0 < N; c0 += chunk) {
    const uint64_t nc = std::min(chunk, N - c0);
    RET(share_strided_on<F>(ctx, ctx->stream, sec + c0, nc, t, n, seed, first_block + c0 * B,
                            tmp.as<E>(), nc, 1));
    RET(transpose_on<E>(ctx, ctx->stream, tmp.as<E>(), n, nc, out + c0 * n));
  }
  return SCLGPU_OK;  // tmp goes back to the pool in stream order
}

// Host pipeline: two streams, chunked; H2D secrets -> share (party-major) ->
// transpose to SCL's [N][n] -> D2H, the copies of one chunk overlapping the
// kernels of the other.
template <class F>
static int share_host(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t t, uint32_t n,
                      const uint8_t seed[16], uint64_t first_block, void* shares) {
  typedef typename F::E E;
  if (!ctx || !seed || ((!secrets || !shares) && N && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n >= (1u << 31)) return fail(ctx, SCLGPU_EINVAL, "n too large");
  CK(cudaSetDevice(ctx->device));
  if (N == 0 || n == 0) return SCLGPU_OK;
  const uint64_t B = ((uint64_t)(t + 1) * F::BYTES + 15) / 16;
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / ((uint64_t)n * sizeof(E)), 1024);
  chunk = std::min(chunk, std::min(N, kHostChunk));
  PoolScope pool_scope(ctx);
  PoolBuf dsec[2], dpm[2], dsm[2];
  const int nbuf = N > chunk ? 2 : 1;
  for (int k = 0; k < nbuf; ++k) {
    CK(dsec[k].alloc(chunk * sizeof(E)));
    CK(dpm[k].alloc(chunk * n * sizeof(E)));
    CK(dsm[k].alloc(chunk * n * sizeof(E)));
  }
  const E* hs = reinterpret_cast<const E*>(secrets);
  E* ho = reinterpret_cast<E*>(shares);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    CK(ctx->stager.h2d(st, dsec[k].p, hs + c0, nc * sizeof(E)));
    RET(share_strided_on<F>(ctx, st, dsec[k].as<E>(), nc, t, n, seed, first_block + c0 * B,
                            dpm[k].as<E>(), nc, 1));
    RET(transpose_on<E>(ctx, st, dpm[k].as<E>(), n, nc, dsm[k].as<E>()));
    CK(ctx->stager.d2h(st, ho + c0 * n, dsm[k].p, nc * n * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}

extern "C" int sclgpu_fp61_shamir_share(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, uint64_t* o) { return guarded(c, [&] { return share_host<F61>(c, s, N, t, n, seed, fb, o); }); }
extern "C" int sclgpu_fp127_shamir_share(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, void* o) { return guarded(c, [&] { return share_host<F127>(c, s, N, t, n, seed, fb, o); }); }
extern "C" int sclgpu_fp61_shamir_share_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, uint64_t* o, int layout) { return guarded(c, [&] { return share_dev<F61>(c, s, N, t, n, seed, fb, o, layout); }); }
extern "C" int sclgpu_fp127_shamir_share_dev(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, void* o, int layout) { return guarded(c, [&] { return share_dev<F127>(c, s, N, t, n, seed, fb, o, layout); }); }

template <class F>
static int share_coeffs_dev(sclgpu_ctx* ctx, const void* d_coeffs, uint64_t N, uint32_t t, uint32_t n,
                            void* d_shares, int layout) {
  typedef typename F::E E;
  if (!ctx || ((!d_coeffs || !d_shares) && N && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n >= (1u << 31)) return fail(ctx, SCLGPU_EINVAL, "n too large");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  CK(cudaSetDevice(ctx->device));
  uint64_t si, sj;
  strides_for(layout, N, n, si, sj);
  return share_coeffs_on<F>(ctx, ctx->stream, (const E*)d_coeffs, N, t, n, (E*)d_shares, si, sj);
}
extern "C" int sclgpu_fp61_shamir_share_coeffs_dev(sclgpu_ctx* c, const uint64_t* k, uint64_t N, uint32_t t, uint32_t n, uint64_t* o, int layout) { return guarded(c, [&] { return share_coeffs_dev<F61>(c, k, N, t, n, o, layout); }); }
extern "C" int sclgpu_fp127_shamir_share_coeffs_dev(sclgpu_ctx* c, const void* k, uint64_t N, uint32_t t, uint32_t n, void* o, int layout) { return guarded(c, [&] { return share_coeffs_dev<F127>(c, k, N, t, n, o, layout); }); }

// ------------------------------------------------------------------ array-valued secrets
// shamirSecretShare on math::Array<FF, W> (shamir.h:52-68 with T = Array; pedersenSecretShare's sharing
// step, pedersen.h:137-138, is W = 2).  The N*W component polynomials are independent, so after the
// keystream is laid out as coefficient planes over the N*W "virtual secrets" (k_expand_coeffs) the
// evaluation is the plain coefficient-plane share on N*W columns, and party-major output
// [n][N][W] is contiguous.  Secret-major [N][n][W] (SCL's N Vectors of Arrays) is one wide transposition.
template <class E>
static int transpose_wide_on(sclgpu_ctx* ctx, cudaStream_t st, const E* d_in, uint64_t rows, uint64_t cols,
                             uint32_t W, E* d_out) {
  if (rows == 0 || cols == 0 || W == 0) return SCLGPU_OK;
  if (W == 1) return transpose_on<E>(ctx, st, d_in, rows, cols, d_out);
  if (sizeof(E) == 8 && W == 2 && ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0)
    return transpose_on<E127>(ctx, st, reinterpret_cast<const E127*>(d_in), rows, cols, reinterpret_cast<E127*>(d_out));  // pairs move as 16-byte elements
  constexpr int CW = sizeof(E) == 8 ? 4 : 2;
  const uint64_t tiles = ((rows + 31) / 32) * ((cols + 31) / 32) * ((W + CW - 1) / CW);
  const int grid = (int)std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * 16);
  k_transpose_wide<E, CW><<<grid, 256, 0, st>>>(d_in, rows, cols, W, d_out);
  CKL();
  return SCLGPU_OK;
}

static bool array_args_ok(sclgpu_ctx* ctx, uint64_t N, uint32_t W, uint32_t n, int& rc) {
  rc = SCLGPU_OK;
  if (W == 0 || W > 4096) rc = fail(ctx, SCLGPU_EINVAL, "array width must be in 1..4096");
  else if (n >= (1u << 31)) rc = fail(ctx, SCLGPU_EINVAL, "n too large");
  else if (N > (~0ull) / W / std::max<uint64_t>(n, 1) / 16) rc = fail(ctx, SCLGPU_EINVAL, "batch too large");
  return rc == SCLGPU_OK;
}

// PRG fused on the tensor-core kernel when the component pairs line up with the keystream blocks
template <class F>
static bool share_array_fused(uint32_t W, uint32_t t, uint32_t n) {
  const bool fits = F::BYTES == 8 ? (t <= kTcMaxT && n <= kTcMaxParties && W % 2 == 0) : (t <= kTcMaxT127 && n <= kTcMaxParties127);
  return fits && share_tc_enabled() && !env_flag("SCLGPU_SHARE_GENERIC");
}

// one chunk of nc sharings, secrets and output on the device; planes: (t+1)*nc*W elements of scratch,
// tmp: n*nc*W elements (secret-major only)
template <class F>
static int share_array_chunk(sclgpu_ctx* ctx, cudaStream_t st, const AesKey& key, uint64_t block0,
                             const typename F::E* d_secrets, uint64_t nc, uint32_t W, uint32_t t, uint32_t n,
                             typename F::E* planes, typename F::E* tmp, typename F::E* d_out, uint64_t out_si,
                             bool secret_major) {
  typedef typename F::E E;
  if (share_array_fused<F>(W, t, n)) {
    const void* d_bmat = nullptr;
    RET(share_tc_bmat<F>(ctx, st, t, n, &d_bmat));
    E* dst = secret_major ? tmp : d_out;
    const uint64_t si = secret_major ? nc * W : out_si;
    ctx->launches++;
    cudaError_t e;
    if constexpr (F::BYTES == 8) {
      e = share61_wide_tc_launch(st, ctx->sm_count, key, ctx->d_t0, d_bmat, block0, d_secrets, nc * W, W, t, n, dst, si, 1);
    } else {
      e = share127_wide_tc_launch(st, ctx->sm_count, key, ctx->d_t0, d_bmat, block0, d_secrets, nc * W, W, t, n, dst, si, 1);
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "launch");
    if (!secret_major) return SCLGPU_OK;
    return transpose_wide_on<E>(ctx, st, tmp, n, nc, W, d_out);
  }
  RET(aes_opt_in(ctx, k_expand_coeffs<F>));
  k_expand_coeffs<F><<<grid_for(ctx, nc, kAesThreads, 1), kAesThreads, kAesDynSmem, st>>>(
      key, ctx->d_t0, block0, d_secrets, nc, t, W, planes);
  CKL();
  if (!secret_major) return share_coeffs_on<F>(ctx, st, planes, nc * W, t, n, d_out, out_si, 1);
  RET(share_coeffs_on<F>(ctx, st, planes, nc * W, t, n, tmp, nc * W, 1));
  return transpose_wide_on<E>(ctx, st, tmp, n, nc, W, d_out);
}

template <class F>
static uint64_t share_array_chunk_len(uint64_t N, uint32_t W, uint32_t t, uint32_t n, uint64_t budget) {
  const uint64_t per = (uint64_t)W * sizeof(typename F::E) * std::max<uint64_t>((uint64_t)t + 1, n);
  uint64_t chunk = std::max<uint64_t>(budget / per, 256);
  return std::min(chunk, N);
}

template <class F>
static int share_array_dev(sclgpu_ctx* ctx, const void* d_secrets, uint64_t N, uint32_t W, uint32_t t, uint32_t n,
                           const uint8_t seed[16], uint64_t first_block, void* d_shares, int layout) {
  typedef typename F::E E;
  if (!ctx || !seed || ((!d_secrets || !d_shares) && N && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  int rc;
  if (!array_args_ok(ctx, N, W, n, rc)) return rc;
  if (W == 1) return share_dev<F>(ctx, d_secrets, N, t, n, seed, first_block, d_shares, layout);  // plain shamirSecretShare
  CK(cudaSetDevice(ctx->device));
  if (N == 0 || n == 0) return SCLGPU_OK;
  const AesKey key = aes_expand(seed);
  const uint64_t B = ((uint64_t)(t + 1) * W * F::BYTES + 15) / 16;
  const bool sm = layout == SCLGPU_SECRET_MAJOR;
  const bool fused = share_array_fused<F>(W, t, n);
  const uint64_t chunk = (fused && !sm) ? N : share_array_chunk_len<F>(N, W, t, n, 512ull << 20);
  StreamBuf planes(ctx, ctx->stream), tmp(ctx, ctx->stream);
  if (!fused) CK(planes.alloc(chunk * W * (uint64_t)(t + 1) * sizeof(E)));
  if (sm) CK(tmp.alloc(chunk * W * (uint64_t)n * sizeof(E)));
  const E* sec = reinterpret_cast<const E*>(d_secrets);
  E* out = reinterpret_cast<E*>(d_shares);
  for (uint64_t c0 = 0; c0 < N; c0 += chunk) {
    const uint64_t nc = std::min(chunk, N - c0);
    RET(share_array_chunk<F>(ctx, ctx->stream, key, first_block + c0 * B, sec + c0 * W, nc, W, t, n, planes.as<E>(),
                             tmp.as<E>(), sm ? out + c0 * n * W : out + c0 * W, N * W, sm));
  }
  return SCLGPU_OK;  // scratch goes back to the pool in stream order
}

template <class F>
static int share_array_host(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t W, uint32_t t, uint32_t n,
                            const uint8_t seed[16], uint64_t first_block, void* shares) {
  typedef typename F::E E;
  if (!ctx || !seed || ((!secrets || !shares) && N && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  int rc;
  if (!array_args_ok(ctx, N, W, n, rc)) return rc;
  if (W == 1) return share_host<F>(ctx, secrets, N, t, n, seed, first_block, shares);  // plain shamirSecretShare
  CK(cudaSetDevice(ctx->device));
  if (N == 0 || n == 0) return SCLGPU_OK;
  const AesKey key = aes_expand(seed);
  const uint64_t B = ((uint64_t)(t + 1) * W * F::BYTES + 15) / 16;
  uint64_t chunk = share_array_chunk_len<F>(N, W, t, n, 256ull << 20);
  chunk = std::min(chunk, kHostChunk);
  PoolScope pool_scope(ctx);
  PoolBuf dsec[2], dpl[2], dpm[2], dsm[2];
  const int nbuf = N > chunk ? 2 : 1;
  for (int k = 0; k < nbuf; ++k) {
    CK(dsec[k].alloc(chunk * W * sizeof(E)));
    if (!share_array_fused<F>(W, t, n)) CK(dpl[k].alloc(chunk * W * (uint64_t)(t + 1) * sizeof(E)));
    CK(dpm[k].alloc(chunk * W * (uint64_t)n * sizeof(E)));
    CK(dsm[k].alloc(chunk * W * (uint64_t)n * sizeof(E)));
  }
  const E* hs = reinterpret_cast<const E*>(secrets);
  E* ho = reinterpret_cast<E*>(shares);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    CK(ctx->stager.h2d(st, dsec[k].p, hs + c0 * W, nc * W * sizeof(E)));
    RET(share_array_chunk<F>(ctx, st, key, first_block + c0 * B, dsec[k].as<E>(), nc, W, t, n, dpl[k].as<E>(),
                             dpm[k].as<E>(), dsm[k].as<E>(), 0, true));
    CK(ctx->stager.d2h(st, ho + c0 * n * W, dsm[k].p, nc * n * W * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}

extern "C" int sclgpu_fp61_shamir_share_array(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t W, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, uint64_t* o) { return guarded(c, [&] { return share_array_host<F61>(c, s, N, W, t, n, seed, fb, o); }); }
extern "C" int sclgpu_fp127_shamir_share_array(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t W, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, void* o) { return guarded(c, [&] { return share_array_host<F127>(c, s, N, W, t, n, seed, fb, o); }); }
extern "C" int sclgpu_fp61_shamir_share_array_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t W, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, uint64_t* o, int layout) { return guarded(c, [&] { return share_array_dev<F61>(c, s, N, W, t, n, seed, fb, o, layout); }); }
extern "C" int sclgpu_fp127_shamir_share_array_dev(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t W, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, void* o, int layout) { return guarded(c, [&] { return share_array_dev<F127>(c, s, N, W, t, n, seed, fb, o, layout); }); }
extern "C" uint64_t sclgpu_share_array_blocks(uint32_t element_bytes, uint32_t W, uint32_t t) { return ((uint64_t)(t + 1) * W * element_bytes + 15) / 16; }

// ------------------------------------------------------------------ per-party packets
// Serializer<math::Vector<FF>>::write (vector.h:596-629 -> serializer.h:160-176): a u32 element count
// (StlVecSizeType, serializer.h:111) followed by the elements' FF::write bytes (ff.h:355-391).  Party i's
// packet is plane i of the device-native layout behind a 4-byte header, so no transposition is needed.
static constexpr uint64_t kPacketHeader = 4;

template <class F>
static int share_packets_host(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t t, uint32_t n,
                              const uint8_t seed[16], uint64_t first_block, uint8_t* const* packets) {
  typedef typename F::E E;
  if (!ctx || !seed || (n && !packets) || (!secrets && N)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n >= (1u << 31)) return fail(ctx, SCLGPU_EINVAL, "n too large");
  if (N > 0xFFFFFFFFull) return fail(ctx, SCLGPU_EINVAL, "a packet holds at most 2^32 - 1 elements");
  CK(cudaSetDevice(ctx->device));
  const uint32_t count = (uint32_t)N;
  for (uint32_t i = 0; i < n; ++i) {
    if (!packets[i]) return fail(ctx, SCLGPU_EINVAL, "null packet buffer");
    std::memcpy(packets[i], &count, kPacketHeader);
  }
  if (N == 0 || n == 0) return SCLGPU_OK;
  const uint64_t B = ((uint64_t)(t + 1) * F::BYTES + 15) / 16;
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / ((uint64_t)n * sizeof(E)), 1024);
  chunk = std::min(chunk, std::min(N, kHostChunk));
  PoolScope pool_scope(ctx);
  PoolBuf dsec[2], dpm[2];
  const int nbuf = N > chunk ? 2 : 1;
  for (int k = 0; k < nbuf; ++k) {
    CK(dsec[k].alloc(chunk * sizeof(E)));
    CK(dpm[k].alloc(chunk * n * sizeof(E)));
  }
  const E* hs = reinterpret_cast<const E*>(secrets);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    CK(ctx->stager.h2d(st, dsec[k].p, hs + c0, nc * sizeof(E)));
    RET(share_strided_on<F>(ctx, st, dsec[k].as<E>(), nc, t, n, seed, first_block + c0 * B, dpm[k].as<E>(), nc, 1));
    for (uint32_t i = 0; i < n; ++i)
      CK(ctx->stager.d2h(st, packets[i] + kPacketHeader + c0 * sizeof(E), dpm[k].as<E>() + (uint64_t)i * nc, nc * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}

template <class F>
static int recover_p_packets_host(sclgpu_ctx* ctx, const uint8_t* const* packets, uint64_t N, uint32_t n,
                                  const void* alphas, const void* x, void* out);

// ------------------------------------------------------------------ additive sharing
template <class F>
static int additive_share_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* d_secrets, uint64_t N, uint32_t n,
                             const uint8_t seed[16], uint64_t first_block, typename F::E* d_out, uint64_t si,
                             uint64_t sj) {
  if (N == 0) return SCLGPU_OK;
  RET(aes_opt_in(ctx, k_additive_share<F>));
  const AesKey key = aes_expand(seed);
  k_additive_share<F><<<grid_for(ctx, N, kAesThreads, 1), kAesThreads, kAesDynSmem, st>>>(key, ctx->d_t0, first_block,
                                                                                        d_secrets, N, n, d_out, si, sj);
  CKL();
  return SCLGPU_OK;
}

template <class F>
static int additive_share_dev(sclgpu_ctx* ctx, const void* d_secrets, uint64_t N, uint32_t n, const uint8_t seed[16],
                              uint64_t first_block, void* d_shares, int layout) {
  typedef typename F::E E;
  if (!ctx || !seed || ((!d_secrets || !d_shares) && N)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n == 0 || n >= (1u << 31)) return fail(ctx, SCLGPU_EINVAL, "additiveShare needs n >= 1");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  CK(cudaSetDevice(ctx->device));
  uint64_t si, sj;
  strides_for(layout, N, n, si, sj);
  return additive_share_on<F>(ctx, ctx->stream, (const E*)d_secrets, N, n, seed, first_block, (E*)d_shares, si, sj);
}

// host pipeline as share_host: chunks on two streams, party-major on the device, transposed to SCL's [N][n]
template <class F>
static int additive_share_host(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t n, const uint8_t seed[16],
                               uint64_t first_block, void* shares) {
  typedef typename F::E E;
  if (!ctx || !seed || ((!secrets || !shares) && N)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n == 0 || n >= (1u << 31)) return fail(ctx, SCLGPU_EINVAL, "additiveShare needs n >= 1");
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / ((uint64_t)n * sizeof(E)), 1024);
  chunk = std::min(chunk, std::min(N, kHostChunk));
  PoolScope pool_scope(ctx);
  PoolBuf dsec[2], dpm[2], dsm[2];
  const int nbuf = N > chunk ? 2 : 1;
  for (int k = 0; k < nbuf; ++k) {
    CK(dsec[k].alloc(chunk * sizeof(E)));
    CK(dpm[k].alloc(chunk * n * sizeof(E)));
    CK(dsm[k].alloc(chunk * n * sizeof(E)));
  }
  const E* hs = reinterpret_cast<const E*>(secrets);
  E* ho = reinterpret_cast<E*>(shares);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    CK(ctx->stager.h2d(st, dsec[k].p, hs + c0, nc * sizeof(E)));
    RET(additive_share_on<F>(ctx, st, dsec[k].as<E>(), nc, n, seed, first_block + c0 * (uint64_t)(n - 1), dpm[k].as<E>(), nc, 1));
    RET(transpose_on<E>(ctx, st, dpm[k].as<E>(), n, nc, dsm[k].as<E>()));
    CK(ctx->stager.d2h(st, ho + c0 * n, dsm[k].p, nc * n * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}

template <class F>
static int additive_recover_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t n, int layout, void* d_out) {
  typedef typename F::E E;
  if (!ctx || (((!d_shares && n) || !d_out) && N)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  uint64_t si, sj;
  strides_for(layout, N, n, si, sj);
  k_additive_recover<F><<<grid_for(ctx, N, 256, 8), 256, 0, ctx->stream>>>((const E*)d_shares, N, n, si, sj, (E*)d_out);
  CKL();
  return SCLGPU_OK;
}

template <class F>
static int additive_recover_host(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n, void* out) {
  typedef typename F::E E;
  if (!ctx || (((!shares && n) || !out) && N)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / (std::max<uint64_t>(n, 1) * sizeof(E)), 1024);
  chunk = std::min(chunk, std::min(N, kHostChunk));
  const int nbuf = N > chunk ? 2 : 1;
  PoolScope pool_scope(ctx);
  PoolBuf dsh[2], dout[2];
  for (int k = 0; k < nbuf; ++k) {
    CK(dsh[k].alloc(chunk * n * sizeof(E)));
    CK(dout[k].alloc(chunk * sizeof(E)));
  }
  const E* hs = reinterpret_cast<const E*>(shares);
  E* ho = reinterpret_cast<E*>(out);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    if (n) CK(ctx->stager.h2d(st, dsh[k].p, hs + c0 * n, nc * n * sizeof(E)));
    k_additive_recover<F><<<grid_for(ctx, nc, 256, 8), 256, 0, st>>>(dsh[k].as<E>(), nc, n, 1, n, dout[k].as<E>());
    CKL();
    CK(ctx->stager.d2h(st, ho + c0, dout[k].p, nc * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}
extern "C" int sclgpu_fp61_additive_share(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, const uint8_t seed[16], uint64_t fb, uint64_t* o) { return guarded(c, [&] { return additive_share_host<F61>(c, s, N, n, seed, fb, o); }); }
extern "C" int sclgpu_fp127_additive_share(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t n, const uint8_t seed[16], uint64_t fb, void* o) { return guarded(c, [&] { return additive_share_host<F127>(c, s, N, n, seed, fb, o); }); }
extern "C" int sclgpu_fp61_additive_share_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, const uint8_t seed[16], uint64_t fb, uint64_t* o, int layout) { return guarded(c, [&] { return additive_share_dev<F61>(c, s, N, n, seed, fb, o, layout); }); }
extern "C" int sclgpu_fp127_additive_share_dev(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t n, const uint8_t seed[16], uint64_t fb, void* o, int layout) { return guarded(c, [&] { return additive_share_dev<F127>(c, s, N, n, seed, fb, o, layout); }); }
extern "C" int sclgpu_fp61_additive_recover(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, uint64_t* o) { return guarded(c, [&] { return additive_recover_host<F61>(c, s, N, n, o); }); }
extern "C" int sclgpu_fp127_additive_recover(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t n, void* o) { return guarded(c, [&] { return additive_recover_host<F127>(c, s, N, n, o); }); }
extern "C" int sclgpu_fp61_additive_recover_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, int layout, uint64_t* o) { return guarded(c, [&] { return additive_recover_dev<F61>(c, s, N, n, layout, o); }); }
extern "C" int sclgpu_fp127_additive_recover_dev(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t n, int layout, void* o) { return guarded(c, [&] { return additive_recover_dev<F127>(c, s, N, n, layout, o); }); }

// ------------------------------------------------------------------ lagrange
template <class F>
static int lagrange_host(sclgpu_ctx* ctx, const void* nodes, uint32_t n, const void* x, void* out) {
  typedef typename F::E E;
  if (!ctx || !x || (!out && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return SCLGPU_OK;
  const E* d_mat = nullptr;
  RET(basis_rows<F>(ctx, ctx->pipe[0], (const E*)nodes, n, (const E*)x, 1, &d_mat));
  CK(cudaMemcpyAsync(out, d_mat, (size_t)n * sizeof(E), cudaMemcpyDeviceToHost, ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  return SCLGPU_OK;
}
extern "C" int sclgpu_fp61_lagrange_basis(sclgpu_ctx* c, const uint64_t* nodes, uint32_t n, const uint64_t* x, uint64_t* o) { return guarded(c, [&] { return lagrange_host<F61>(c, nodes, n, x, o); }); }
extern "C" int sclgpu_fp127_lagrange_basis(sclgpu_ctx* c, const void* nodes, uint32_t n, const void* x, void* o) { return guarded(c, [&] { return lagrange_host<F127>(c, nodes, n, x, o); }); }

// Matrix::hyperInvertible(n, m), matrix.h:462-475: row i = computeLagrangeBasis(range(1, m+1), -i); the
// int overload (lagrange.h:80-82) makes -i the field element p - i (FF(int), mersenne61.cc:38-40).
template <class F>
static int hyper_invertible_host(sclgpu_ctx* ctx, uint32_t n, uint32_t m, void* out) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (n == 0 || m == 0) return fail(ctx, SCLGPU_EINVAL, "n or m cannot be 0");  // matrix.h:165
  if (!out) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n >= (1u << 31) || m >= (1u << 31)) return fail(ctx, SCLGPU_EINVAL, "n or m too large");
  CK(cudaSetDevice(ctx->device));
  std::vector<E> xs(n);
  for (uint32_t i = 0; i < n; ++i) xs[i] = F::neg(F::from_u32(i));
  const E* d_mat = nullptr;
  RET(basis_rows<F>(ctx, ctx->pipe[0], nullptr, m, xs.data(), n, &d_mat));
  CK(cudaMemcpyAsync(out, d_mat, (size_t)n * m * sizeof(E), cudaMemcpyDeviceToHost, ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  return SCLGPU_OK;
}
extern "C" int sclgpu_fp61_hyper_invertible(sclgpu_ctx* c, uint32_t n, uint32_t m, uint64_t* o) { return guarded(c, [&] { return hyper_invertible_host<F61>(c, n, m, o); }); }
extern "C" int sclgpu_fp127_hyper_invertible(sclgpu_ctx* c, uint32_t n, uint32_t m, void* o) { return guarded(c, [&] { return hyper_invertible_host<F127>(c, n, m, o); }); }

// ------------------------------------------------------------------ recover P
template <class F>
static int recover_p_basis(sclgpu_ctx* ctx, cudaStream_t st, uint32_t n, const void* alphas, const void* x,
                           const typename F::E** d_basis) {
  typedef typename F::E E;
  E xx = F::zero();
  if (alphas != nullptr) {
    if (!x) return fail(ctx, SCLGPU_EINVAL, "x is required with explicit alphas");
    xx = *reinterpret_cast<const E*>(x);
  }
  return basis_rows<F>(ctx, st, (const E*)alphas, n, &xx, 1, d_basis);
}

template <class F>
static int recover_p_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t n, int layout,
                         const void* alphas, const void* x, void* d_out) {
  typedef typename F::E E;
  if (!ctx || ((!d_shares && n) || !d_out) && N) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  const E* d_basis = nullptr;
  RET(recover_p_basis<F>(ctx, ctx->stream, n, alphas, x, &d_basis));
  uint64_t si, sj;
  strides_for(layout, N, n, si, sj);
  return recover_p_on<F>(ctx, ctx->stream, (const E*)d_shares, N, n, si, sj, d_basis, (E*)d_out);
}

template <class F>
static int recover_p_host(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n, const void* alphas,
                          const void* x, void* out) {
  typedef typename F::E E;
  if (!ctx || ((!shares && n) || !out) && N) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  const E* d_basis = nullptr;
  RET(recover_p_basis<F>(ctx, ctx->pipe[0], n, alphas, x, &d_basis));
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / (std::max<uint64_t>(n, 1) * sizeof(E)), 1024);
  chunk = std::min(chunk, std::min(N, kHostChunk));
  chunk &= ~1ull;  // even chunks: 128-bit loads in the plane kernel
  if (chunk == 0) chunk = N;
  const int nbuf = N > chunk ? 2 : 1;
  PoolScope pool_scope(ctx);
  PoolBuf dsh[2], dpm[2], dout[2];
  for (int k = 0; k < nbuf; ++k) {
    CK(dsh[k].alloc(chunk * n * sizeof(E)));
    CK(dpm[k].alloc(chunk * n * sizeof(E)));
    CK(dout[k].alloc(chunk * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  const E* hs = reinterpret_cast<const E*>(shares);
  E* ho = reinterpret_cast<E*>(out);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    if (n) CK(ctx->stager.h2d(st, dsh[k].p, hs + c0 * n, nc * n * sizeof(E)));
    // SCL's [N][n] -> party-major planes on the device (coalesced on both sides), then the plane kernel
    RET(transpose_on<E>(ctx, st, dsh[k].as<E>(), nc, n, dpm[k].as<E>()));
    RET(recover_p_on<F>(ctx, st, dpm[k].as<E>(), nc, n, nc, 1, d_basis, dout[k].as<E>()));
    CK(ctx->stager.d2h(st, ho + c0, dout[k].p, nc * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}
// shamirRecoverP from the n packets a reconstructing party received (packet i = Vector of party i's
// shares of all N secrets): the planes go to the device as they are, no transposition.
template <class F>
static int recover_p_packets_host(sclgpu_ctx* ctx, const uint8_t* const* packets, uint64_t N, uint32_t n,
                                  const void* alphas, const void* x, void* out) {
  typedef typename F::E E;
  if (!ctx || (n && !packets) || (!out && N)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (N > 0xFFFFFFFFull) return fail(ctx, SCLGPU_EINVAL, "a packet holds at most 2^32 - 1 elements");
  for (uint32_t i = 0; i < n; ++i) {
    if (!packets[i]) return fail(ctx, SCLGPU_EINVAL, "null packet buffer");
    uint32_t count;
    std::memcpy(&count, packets[i], kPacketHeader);
    if (count != (uint32_t)N) return fail(ctx, SCLGPU_EINVAL, "Vec sizes mismatch");  // vector.h:483
  }
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  const E* d_basis = nullptr;
  RET(recover_p_basis<F>(ctx, ctx->pipe[0], n, alphas, x, &d_basis));
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / (std::max<uint64_t>(n, 1) * sizeof(E)), 1024);
  chunk = std::min(chunk, std::min(N, kHostChunk)) & ~1ull;  // even: 128-bit loads in the plane kernel
  if (chunk == 0) chunk = N;
  const int nbuf = N > chunk ? 2 : 1;
  PoolScope pool_scope(ctx);
  PoolBuf dsh[2], dout[2];
  for (int k = 0; k < nbuf; ++k) {
    CK(dsh[k].alloc(chunk * std::max<uint64_t>(n, 1) * sizeof(E)));
    CK(dout[k].alloc(chunk * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  E* ho = reinterpret_cast<E*>(out);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    for (uint32_t i = 0; i < n; ++i)
      CK(ctx->stager.h2d(st, dsh[k].as<E>() + (uint64_t)i * nc, packets[i] + kPacketHeader + c0 * sizeof(E), nc * sizeof(E)));
    // wire bytes from other parties: Serializer<Vector<FF>>::read goes through FF::read, i.e. `% p` on every word
    // (vector.h:623-626, ff.h:63-67, mersenne61.cc:87-90) -- done here in place, so that words in [p, 2^64) /
    // [p, 2^128) reconstruct to what the reference reconstructs
    RET(from_bytes_on<F>(ctx, st, reinterpret_cast<const uint8_t*>(dsh[k].p), (uint64_t)n * nc, dsh[k].as<E>()));
    RET(recover_p_on<F>(ctx, st, dsh[k].as<E>(), nc, n, nc, 1, d_basis, dout[k].as<E>()));
    CK(ctx->stager.d2h(st, ho + c0, dout[k].p, nc * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}
extern "C" int sclgpu_fp61_shamir_share_packets(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, uint8_t* const* p) { return guarded(c, [&] { return share_packets_host<F61>(c, s, N, t, n, seed, fb, p); }); }
extern "C" int sclgpu_fp127_shamir_share_packets(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, uint8_t* const* p) { return guarded(c, [&] { return share_packets_host<F127>(c, s, N, t, n, seed, fb, p); }); }
extern "C" int sclgpu_fp61_recover_p_packets(sclgpu_ctx* c, const uint8_t* const* p, uint64_t N, uint32_t n, const uint64_t* a, const uint64_t* x, uint64_t* o) { return guarded(c, [&] { return recover_p_packets_host<F61>(c, p, N, n, a, x, o); }); }
extern "C" int sclgpu_fp127_recover_p_packets(sclgpu_ctx* c, const uint8_t* const* p, uint64_t N, uint32_t n, const void* a, const void* x, void* o) { return guarded(c, [&] { return recover_p_packets_host<F127>(c, p, N, n, a, x, o); }); }
extern "C" uint64_t sclgpu_packet_bytes(uint32_t element_bytes, uint64_t n_elements) { return kPacketHeader + (uint64_t)element_bytes * n_elements; }
extern "C" int sclgpu_fp61_recover_p(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, const uint64_t* a, const uint64_t* x, uint64_t* o) { return guarded(c, [&] { return recover_p_host<F61>(c, s, N, n, a, x, o); }); }
extern "C" int sclgpu_fp127_recover_p(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t n, const void* a, const void* x, void* o) { return guarded(c, [&] { return recover_p_host<F127>(c, s, N, n, a, x, o); }); }
extern "C" int sclgpu_fp61_recover_p_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, int layout, const uint64_t* a, const uint64_t* x, uint64_t* o) { return guarded(c, [&] { return recover_p_dev<F61>(c, s, N, n, layout, a, x, o); }); }
extern "C" int sclgpu_fp127_recover_p_dev(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t n, int layout, const void* a, const void* x, void* o) { return guarded(c, [&] { return recover_p_dev<F127>(c, s, N, n, layout, a, x, o); }); }

// ------------------------------------------------------------------ recover P + all-gather over peer memory
// shamirRecoverP of this rank's slice of a batch, the result written straight into EVERY rank's copy of the
// gathered vector ("gather reconstructed values"): d_dsts[r] = base of rank r's gathered buffer as
// mapped on this device (its own cudaMalloc memory for r = self, peer memory for the others, sclgpu_ipc_open /
// sclgpu_enable_peer), and element j of the slice goes to d_dsts[r][offset + j].  One kernel: the NVLink stores
// are posted while the planes are still being read -- no separate collective, no staging copy.
static int recover_p_gather61_dev(sclgpu_ctx* ctx, const uint64_t* d_shares, uint64_t N, uint32_t n, const uint64_t* alphas,
                                  const uint64_t* x, uint64_t* const* d_dsts, uint32_t n_dsts, uint64_t offset) {
  if (!ctx || (!d_shares && n && N) || !d_dsts) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n_dsts < 1 || n_dsts > 8) return fail(ctx, SCLGPU_EINVAL, "1..8 gather destinations");
  for (uint32_t r = 0; r < n_dsts; ++r)
    if (!d_dsts[r]) return fail(ctx, SCLGPU_EINVAL, "null gather destination");
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  const uint64_t* d_basis = nullptr;
  RET(recover_p_basis<F61>(ctx, ctx->stream, n, alphas, x, &d_basis));
  GatherDst gd;
  std::memset(&gd, 0, sizeof(gd));
  gd.count = n_dsts;
  for (uint32_t r = 0; r < n_dsts; ++r) gd.dst[r] = d_dsts[r] + offset;
  return recover_p_on<F61>(ctx, ctx->stream, d_shares, N, n, N, 1, d_basis, gd.dst[0], &gd);
}
extern "C" int sclgpu_fp61_recover_p_gather_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, const uint64_t* a, const uint64_t* x, uint64_t* const* d, uint32_t nd, uint64_t off) { return guarded(c, [&] { return recover_p_gather61_dev(c, s, N, n, a, x, d, nd, off); }); }

// Peer memory plumbing for the call above.  One process per GPU: export the handle of a sclgpu_malloc'ed buffer,
// pass the 64 bytes to the other ranks by any means (torch.distributed, MPI, a socket), open it there.
extern "C" int sclgpu_ipc_export(sclgpu_ctx* ctx, void* d_ptr, uint8_t handle[64]) {
  if (!ctx || !d_ptr || !handle) return fail(ctx, SCLGPU_EINVAL, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, d_ptr));
  std::memcpy(handle, &h, 64);
  return SCLGPU_OK;
}
extern "C" int sclgpu_ipc_open(sclgpu_ctx* ctx, const uint8_t handle[64], void** d_ptr) {
  if (!ctx || !d_ptr || !handle) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, 64);
  CK(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return SCLGPU_OK;
}
extern "C" int sclgpu_ipc_close(sclgpu_ctx* ctx, void* d_ptr) {
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  CK(cudaIpcCloseMemHandle(d_ptr));
  return SCLGPU_OK;
}
// Single process driving several GPUs: let this context's device address memory of `peer_device` directly.
extern "C" int sclgpu_enable_peer(sclgpu_ctx* ctx, int peer_device) {
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  if (peer_device == ctx->device) return SCLGPU_OK;
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, ctx->device, peer_device));
  if (!can) return fail(ctx, SCLGPU_ECUDA, "devices cannot access each other's memory");
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) {
    cudaGetLastError();
    return SCLGPU_OK;
  }
  if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaDeviceEnablePeerAccess");
  return SCLGPU_OK;
}

// ------------------------------------------------------------------ share + recover P in one launch
// N x { shamirSecretShare (shamir.h:52-68), shamirRecoverP (shamir.h:82-104) } on party-major planes.  With
// d_rec_shares == d_shares the sharings produced by this call are reconstructed (the round trip of one batch);
// otherwise d_rec_shares is another batch of N sharings, reconstructed under the share work of this one.
// Shapes the fused kernel does not take (t > 15, n > 32, another field) run as the two kernels back to back.
static int share_recover61_dev(sclgpu_ctx* ctx, const uint64_t* d_secrets, uint64_t N, uint32_t t, uint32_t n,
                               const uint8_t seed[16], uint64_t first_block, uint64_t* d_shares,
                               const uint64_t* d_rec_shares, const uint64_t* alphas, const uint64_t* x,
                               uint64_t* d_out, uint64_t* const* d_dsts = nullptr, uint32_t n_dsts = 0, uint64_t offset = 0) {
  if (!ctx || !seed || ((!d_secrets || !d_shares || !d_rec_shares) && N && n))
    return fail(ctx, SCLGPU_EINVAL, "null argument");
  GatherDst gd;
  std::memset(&gd, 0, sizeof(gd));
  if (d_dsts) {  // reconstructed secrets gathered into every destination at `offset` instead of d_out
    if (n_dsts < 1 || n_dsts > 8) return fail(ctx, SCLGPU_EINVAL, "1..8 gather destinations");
    for (uint32_t r = 0; r < n_dsts; ++r) {
      if (!d_dsts[r]) return fail(ctx, SCLGPU_EINVAL, "null gather destination");
      gd.dst[r] = d_dsts[r] + offset;
    }
    gd.count = n_dsts;
    d_out = gd.dst[0];
  }
  if (!d_out && N && n) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n >= (1u << 31)) return fail(ctx, SCLGPU_EINVAL, "n too large");
  CK(cudaSetDevice(ctx->device));
  if (N == 0 || n == 0) return SCLGPU_OK;
  cudaStream_t st = ctx->stream;
  const uint64_t* d_basis = nullptr;
  RET(recover_p_basis<F61>(ctx, st, n, alphas, x, &d_basis));
  // the reconstruction warps use 128-bit accesses (two secrets per thread): even N, 16-byte aligned planes and output
  const bool fused = t <= kTcMaxT && n <= kTcMaxParties && share_tc_enabled() && !env_flag("SCLGPU_SHARE_GENERIC") &&
                     !env_flag("SCLGPU_NO_FUSED_STEP") && N % 2 == 0 &&
                     ((reinterpret_cast<uintptr_t>(d_rec_shares) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
  uintptr_t galign = 0;
  for (uint32_t r = 0; r < gd.count; ++r) galign |= reinterpret_cast<uintptr_t>(gd.dst[r]);
  if (!fused || (galign & 15)) {
    RET(share_strided_on<F61>(ctx, st, d_secrets, N, t, n, seed, first_block, d_shares, N, 1));
    return recover_p_on<F61>(ctx, st, d_rec_shares, N, n, N, 1, d_basis, d_out, gd.count ? &gd : nullptr);
  }
  auto it = ctx->rec_basis_cache.find(d_basis);
  if (it == ctx->rec_basis_cache.end()) {
    uint64_t hb[kTcMaxParties];
    CK(cudaMemcpyAsync(hb, d_basis, (size_t)n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    RecBasis61 rb;
    std::memset(&rb, 0, sizeof(rb));
    for (uint32_t i = 0; i < n; ++i) {
      rb.l[i][0] = (uint32_t)(hb[i] & 0x1FFFFFu);
      rb.l[i][1] = (uint32_t)((hb[i] >> 21) & 0x1FFFFFu);
      rb.l[i][2] = (uint32_t)(hb[i] >> 42);
    }
    it = ctx->rec_basis_cache.emplace(d_basis, rb).first;
  }
  if (!ctx->sr_prepared) {
    CK(share_recover61_prepare());
    ctx->sr_prepared = true;
  }
  const void* d_bmat = nullptr;
  RET(share_tc_bmat<F61>(ctx, st, t, n, &d_bmat));
  const AesKey key = aes_expand(seed);
  ctx->launches++;
  cudaError_t e = share_recover61_launch(st, ctx->sm_count, env_int("SCLGPU_SR_WARPS", 4), key, it->second, ctx->d_t0, d_bmat, first_block, d_secrets, N, t, n,
                                         d_shares, d_rec_shares, d_out, gd.count ? &gd : nullptr);
  if (e != cudaSuccess) return cuda_fail(ctx, e, "launch");
  return SCLGPU_OK;
}
extern "C" int sclgpu_fp61_shamir_share_recover_gather_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, uint64_t* sh, const uint64_t* rs, const uint64_t* a, const uint64_t* x, uint64_t* const* d, uint32_t nd, uint64_t off) { return guarded(c, [&] { return d ? share_recover61_dev(c, s, N, t, n, seed, fb, sh, rs, a, x, nullptr, d, nd, off) : fail(c, SCLGPU_EINVAL, "null argument"); }); }
extern "C" int sclgpu_fp61_shamir_share_recover_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t t, uint32_t n, const uint8_t seed[16], uint64_t fb, uint64_t* sh, const uint64_t* rs, const uint64_t* a, const uint64_t* x, uint64_t* o) { return guarded(c, [&] { return share_recover61_dev(c, s, N, t, n, seed, fb, sh, rs, a, x, o); }); }

// shamirRecoverP on Vector<Array<FF, W>> (shamir.h:100-104 with T = Array): the basis of nodes 1..n at 0
// applied component-wise, i.e. the plane kernel on N*W columns.
template <class F>
static int recover_p_array_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t W, uint32_t n, int layout,
                               void* d_out) {
  typedef typename F::E E;
  if (!ctx || ((!d_shares && n) || !d_out) && N) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  int rc;
  if (!array_args_ok(ctx, N, W, n, rc)) return rc;
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  const E* d_basis = nullptr;
  RET(recover_p_basis<F>(ctx, ctx->stream, n, nullptr, nullptr, &d_basis));
  const E* in = reinterpret_cast<const E*>(d_shares);
  E* out = reinterpret_cast<E*>(d_out);
  if (layout == SCLGPU_PARTY_MAJOR || n == 0)
    return recover_p_on<F>(ctx, ctx->stream, in, N * W, n, N * W, 1, d_basis, out);
  uint64_t chunk = std::max<uint64_t>((512ull << 20) / ((uint64_t)n * W * sizeof(E)), 256);
  chunk = std::min(chunk, N);
  StreamBuf tmp(ctx, ctx->stream);
  CK(tmp.alloc(chunk * n * W * sizeof(E)));
  for (uint64_t c0 = 0; c0 < N; c0 += chunk) {
    const uint64_t nc = std::min(chunk, N - c0);
    RET(transpose_wide_on<E>(ctx, ctx->stream, in + c0 * n * W, nc, n, W, tmp.as<E>()));
    RET(recover_p_on<F>(ctx, ctx->stream, tmp.as<E>(), nc * W, n, nc * W, 1, d_basis, out + c0 * W));
  }
  return SCLGPU_OK;
}

template <class F>
static int recover_p_array_host(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t W, uint32_t n, void* out) {
  typedef typename F::E E;
  if (!ctx || ((!shares && n) || !out) && N) return fail(ctx, SCLGPU_EINVAL, "null argument");
  int rc;
  if (!array_args_ok(ctx, N, W, n, rc)) return rc;
  CK(cudaSetDevice(ctx->device));
  if (N == 0) return SCLGPU_OK;
  const E* d_basis = nullptr;
  RET(recover_p_basis<F>(ctx, ctx->pipe[0], n, nullptr, nullptr, &d_basis));
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / (std::max<uint64_t>(n, 1) * W * sizeof(E)), 256);
  chunk = std::min(chunk, std::min(N, kHostChunk));
  const int nbuf = N > chunk ? 2 : 1;
  PoolScope pool_scope(ctx);
  PoolBuf dsh[2], dpm[2], dout[2];
  for (int k = 0; k < nbuf; ++k) {
    CK(dsh[k].alloc(chunk * n * W * sizeof(E)));
    CK(dpm[k].alloc(chunk * n * W * sizeof(E)));
    CK(dout[k].alloc(chunk * W * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  const E* hs = reinterpret_cast<const E*>(shares);
  E* ho = reinterpret_cast<E*>(out);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    if (n) CK(ctx->stager.h2d(st, dsh[k].p, hs + c0 * n * W, nc * n * W * sizeof(E)));
    RET(transpose_wide_on<E>(ctx, st, dsh[k].as<E>(), nc, n, W, dpm[k].as<E>()));
    RET(recover_p_on<F>(ctx, st, dpm[k].as<E>(), nc * W, n, nc * W, 1, d_basis, dout[k].as<E>()));
    CK(ctx->stager.d2h(st, ho + c0 * W, dout[k].p, nc * W * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}
extern "C" int sclgpu_fp61_recover_p_array(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t W, uint32_t n, uint64_t* o) { return guarded(c, [&] { return recover_p_array_host<F61>(c, s, N, W, n, o); }); }
extern "C" int sclgpu_fp127_recover_p_array(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t W, uint32_t n, void* o) { return guarded(c, [&] { return recover_p_array_host<F127>(c, s, N, W, n, o); }); }
extern "C" int sclgpu_fp61_recover_p_array_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t W, uint32_t n, int layout, uint64_t* o) { return guarded(c, [&] { return recover_p_array_dev<F61>(c, s, N, W, n, layout, o); }); }
extern "C" int sclgpu_fp127_recover_p_array_dev(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t W, uint32_t n, int layout, void* o) { return guarded(c, [&] { return recover_p_array_dev<F127>(c, s, N, W, n, layout, o); }); }

// ------------------------------------------------------------------ recover D
static int finish_detect(sclgpu_ctx* ctx, cudaStream_t st, uint64_t* n_detected) {
  unsigned long long bad = 0;
  CK(cudaMemcpyAsync(&bad, ctx->d_count, sizeof(bad), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (n_detected) *n_detected = bad;
  if (bad) return fail(ctx, SCLGPU_EDETECT, "error detected during recovery");
  return SCLGPU_OK;
}

template <class F>
static int recover_d_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t n_given, int layout,
                         uint32_t t, const void* alphas, uint32_t n_alphas, uint32_t d, const void* x,
                         void* d_out, uint8_t* d_err, uint64_t* n_detected) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  CK(cudaSetDevice(ctx->device));
  uint32_t m = 0, n_checks = 0;
  const E* d_mat = nullptr;
  RET(recover_d_matrix<F>(ctx, ctx->stream, n_given, t, (const E*)alphas, n_alphas, d, (const E*)x, m,
                          n_checks, &d_mat));
  if (n_detected) *n_detected = 0;
  if (N == 0) return SCLGPU_OK;
  if (!d_shares || !d_out || !d_err) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaMemsetAsync(ctx->d_count, 0, sizeof(unsigned long long), ctx->stream));
  uint64_t si, sj;
  strides_for(layout, N, n_given, si, sj);
  RET(recover_d_on<F>(ctx, ctx->stream, (const E*)d_shares, N, si, sj, m, n_checks, d_mat, (E*)d_out, d_err));
  return finish_detect(ctx, ctx->stream, n_detected);
}

template <class F>
static int recover_d_host(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n_given, uint32_t t,
                          const void* alphas, uint32_t n_alphas, uint32_t d, const void* x, void* out,
                          uint8_t* err, uint64_t* n_detected) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  CK(cudaSetDevice(ctx->device));
  uint32_t m = 0, n_checks = 0;
  const E* d_mat = nullptr;
  RET(recover_d_matrix<F>(ctx, ctx->pipe[0], n_given, t, (const E*)alphas, n_alphas, d, (const E*)x, m,
                          n_checks, &d_mat));
  if (n_detected) *n_detected = 0;
  if (N == 0) return SCLGPU_OK;
  if (!shares || !out || !err) return fail(ctx, SCLGPU_EINVAL, "null argument");
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / (std::max<uint64_t>(n_given, 1) * sizeof(E)), 1024);
  chunk = std::min(chunk, std::min(N, kHostChunk));
  const int nbuf = N > chunk ? 2 : 1;
  PoolScope pool_scope(ctx);
  PoolBuf dsh[2], dout[2], derr[2];
  for (int k = 0; k < nbuf; ++k) {
    CK(dsh[k].alloc(chunk * n_given * sizeof(E)));
    CK(dout[k].alloc(chunk * sizeof(E)));
    CK(derr[k].alloc(chunk));
  }
  CK(cudaMemsetAsync(ctx->d_count, 0, sizeof(unsigned long long), ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  const E* hs = reinterpret_cast<const E*>(shares);
  E* ho = reinterpret_cast<E*>(out);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    CK(ctx->stager.h2d(st, dsh[k].p, hs + c0 * n_given, nc * n_given * sizeof(E)));
    RET(recover_d_on<F>(ctx, st, dsh[k].as<E>(), nc, 1, n_given, m, n_checks, d_mat, dout[k].as<E>(),
                        derr[k].as<uint8_t>()));
    CK(ctx->stager.d2h(st, ho + c0, dout[k].p, nc * sizeof(E)));
    CK(ctx->stager.d2h(st, err + c0, derr[k].p, nc));
  }
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  const int rc = finish_detect(ctx, ctx->pipe[0], n_detected);
  CK(ctx->stager.drain());
  return rc;
}
extern "C" int sclgpu_fp61_recover_d(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t ng, uint32_t t, const uint64_t* a, uint32_t na, uint32_t d, const uint64_t* x, uint64_t* o, uint8_t* e, uint64_t* nd) { return guarded(c, [&] { return recover_d_host<F61>(c, s, N, ng, t, a, na, d, x, o, e, nd); }); }
extern "C" int sclgpu_fp127_recover_d(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t ng, uint32_t t, const void* a, uint32_t na, uint32_t d, const void* x, void* o, uint8_t* e, uint64_t* nd) { return guarded(c, [&] { return recover_d_host<F127>(c, s, N, ng, t, a, na, d, x, o, e, nd); }); }
extern "C" int sclgpu_fp61_recover_d_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t ng, int layout, uint32_t t, const uint64_t* a, uint32_t na, uint32_t d, const uint64_t* x, uint64_t* o, uint8_t* e, uint64_t* nd) { return guarded(c, [&] { return recover_d_dev<F61>(c, s, N, ng, layout, t, a, na, d, x, o, e, nd); }); }
extern "C" int sclgpu_fp127_recover_d_dev(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t ng, int layout, uint32_t t, const void* a, uint32_t na, uint32_t d, const void* x, void* o, uint8_t* e, uint64_t* nd) { return guarded(c, [&] { return recover_d_dev<F127>(c, s, N, ng, layout, t, a, na, d, x, o, e, nd); }); }

// ------------------------------------------------------------------ recover C
// alphas stay HOST pointers (n values); d_* are device pointers
template <class F>
static int recover_c_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* d_shares, uint64_t N, uint32_t n,
                        uint64_t si, uint64_t sj, const typename F::E* alphas, typename F::E* d_f,
                        typename F::E* d_e, uint8_t* d_status, uint64_t* n_failed) {
  typedef typename F::E E;
  if (n == 0) return fail(ctx, SCLGPU_EINVAL, "shamirRecoverC needs at least one share");
  const uint32_t t = (n - 1) / 3, np = 3 * t + 1;
  // 3t+1 <= 32: a warp per sharing (k_recover_c); larger: a CTA per sharing (k_recover_c_cta), its system in shared memory
  const bool big = np > 32;
  if (big && (np > kRecoverCBigMaxPoints || recover_c_big_smem<F>(np) > 227 * 1024))
    return fail(ctx, SCLGPU_EINVAL, "recover_c: the (3t+1) x (3t+2) system exceeds shared memory (Fp61: n <= 166, Fp127: n <= 118)");
  if (n_failed) *n_failed = 0;
  if (N == 0) return SCLGPU_OK;
  std::vector<E> al(np);
  for (uint32_t i = 0; i < np; ++i) al[i] = alphas ? alphas[i] : F::from_u32(i + 1);
  DevBuf dal;
  CK(dal.alloc(np * sizeof(E)));
  CK(cudaMemcpyAsync(dal.p, al.data(), np * sizeof(E), cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(ctx->d_count, 0, sizeof(unsigned long long), st));
  // Error-free sharings first (k_recover_c_clean, one thread each); only the others need the elimination.
  // Needs pairwise distinct nodes (else the reference's own behaviour is the e = 0 oddity handled by k_recover_c).
  bool distinct = true;
  for (uint32_t i = 0; i < np && distinct; ++i)
    for (uint32_t j = i + 1; j < np; ++j)
      if (F::eq(al[i], al[j])) distinct = false;
  StreamBuf dpending(ctx, st);
  DevBuf dcoef;
  uint32_t* d_pending = nullptr;
  unsigned long long* d_n_pending = nullptr;
  std::vector<E> coef;
  if (distinct && N < (1ull << 32) && !env_flag("SCLGPU_RECOVER_C_FULL")) {
    const uint32_t m = t + 1;
    const E* d_check = nullptr;
    if (t > 0) RET(basis_rows<F>(ctx, st, al.data(), m, al.data() + m, 2 * t, &d_check));
    // coef[k][i] = coefficient of x^k in the Lagrange basis polynomial of node a_i among a_0..a_t
    coef.assign((size_t)m * m, F::zero());
    std::vector<E> poly(m + 1);
    for (uint32_t i = 0; i < m; ++i) {
      std::fill(poly.begin(), poly.end(), F::zero());
      poly[0] = F::one();
      uint32_t deg = 0;
      E den = F::one();
      for (uint32_t j = 0; j < m; ++j) {
        if (j == i) continue;
        for (uint32_t d = deg + 1; d >= 1; --d) poly[d] = F::sub(poly[d - 1], F::mul(al[j], poly[d]));
        poly[0] = F::neg(F::mul(al[j], poly[0]));
        ++deg;
        den = F::mul(den, F::sub(al[i], al[j]));
      }
      const E inv = F::inv(den);
      for (uint32_t k = 0; k < m; ++k) coef[(size_t)k * m + i] = F::mul(poly[k], inv);
    }
    CK(dcoef.alloc(coef.size() * sizeof(E)));
    CK(cudaMemcpyAsync(dcoef.p, coef.data(), coef.size() * sizeof(E), cudaMemcpyHostToDevice, st));
    // stream-ordered scratch: stays cached in the device's pool between calls (a cudaMalloc / cudaFree pair
    // would cost milliseconds and a device synchronisation per call)
    CK(dpending.alloc(N * sizeof(uint32_t) + sizeof(unsigned long long)));
    d_n_pending = dpending.as<unsigned long long>();  // counter first (8-byte aligned), then the index list
    d_pending = reinterpret_cast<uint32_t*>(d_n_pending + 1);
    CK(cudaMemsetAsync(d_n_pending, 0, sizeof(unsigned long long), st));
    if (t <= kRecoverCMaxT) {
      const size_t csm = (size_t)(3 * t + 1) * m * sizeof(E);
      k_recover_c_clean<F><<<grid_for(ctx, N, 256, 4), 256, csm, st>>>(d_shares, N, si, sj, t, d_check, dcoef.as<E>(), d_f,
                                                                      d_e, d_status, d_pending, d_n_pending);
    } else {
      k_recover_c_clean_any<F><<<grid_for(ctx, N, 256, 4), 256, 0, st>>>(d_shares, N, si, sj, t, d_check, dcoef.as<E>(), d_f,
                                                                        d_e, d_status, d_pending, d_n_pending);
    }
    CKL();
  }
  const int quick = (distinct && !env_flag("SCLGPU_RECOVER_C_FULL")) ? 1 : 0;
  if (big) {
    const size_t bsm = recover_c_big_smem<F>(np);
    CK(cudaFuncSetAttribute(k_recover_c_cta<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
    const int threads = (int)((np + 31u) / 32u * 32u);
    const int per_sm = std::max(1, std::min((int)((227 * 1024) / bsm), 2048 / threads));
    const int grid = (int)std::min<uint64_t>(N, (uint64_t)ctx->sm_count * per_sm);
    k_recover_c_cta<F><<<grid, threads, bsm, st>>>(d_shares, N, si, sj, t, dal.as<E>(), d_f, d_e, d_status, ctx->d_count,
                                                  d_pending, d_n_pending, quick);
    CKL();
    unsigned long long bad_big = 0;
    CK(cudaMemcpyAsync(&bad_big, ctx->d_count, sizeof(bad_big), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));  // also keeps `dal` alive until the kernel is done
    if (n_failed) *n_failed = bad_big;
    if (bad_big) return fail(ctx, SCLGPU_ECORRECT, "could not correct shares");
    return SCLGPU_OK;
  }
  // Sharings with errors: syndrome decoding, one thread each (k_recover_c_syndrome); what it cannot settle -- more than
  // t errors -- goes on, compacted again, to the elimination kernel below.
  StreamBuf dpending2(ctx, st);
  DevBuf dsyn;
  if (d_pending != nullptr && t >= 1 && t <= kSynMaxT && !env_flag("SCLGPU_RECOVER_C_NOSYN")) {
    const uint32_t m = t + 1;
    std::vector<E> cst((size_t)3 * np + (size_t)m * m);
    for (uint32_t i = 0; i < np; ++i) {
      E prod = F::one();
      for (uint32_t j = 0; j < np; ++j)
        if (j != i) prod = F::mul(prod, F::sub(al[i], al[j]));
      cst[i] = al[i];
      cst[np + i] = F::inv(prod);
      cst[2 * np + i] = prod;
    }
    std::copy(coef.begin(), coef.end(), cst.begin() + 3 * np);
    CK(dsyn.alloc(cst.size() * sizeof(E)));
    CK(cudaMemcpyAsync(dsyn.p, cst.data(), cst.size() * sizeof(E), cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));  // cst is a local
    CK(dpending2.alloc(N * sizeof(uint32_t) + sizeof(unsigned long long)));
    unsigned long long* d_n2 = dpending2.as<unsigned long long>();
    uint32_t* d_p2 = reinterpret_cast<uint32_t*>(d_n2 + 1);
    CK(cudaMemsetAsync(d_n2, 0, sizeof(unsigned long long), st));
    const size_t ssm = cst.size() * sizeof(E);
    k_recover_c_syndrome<F><<<grid_for(ctx, N, 128, 8), 128, ssm, st>>>(d_shares, si, sj, t, dsyn.as<E>(), d_f, d_e, d_status,
                                                                       d_pending, d_n_pending, d_p2, d_n2);
    CKL();
    d_pending = d_p2;
    d_n_pending = d_n2;
  }
  const int warps_per_cta = 8;
  const size_t smem = (size_t)warps_per_cta * ((size_t)np * (np + 1) + 3 * np) * sizeof(E);
  CK(cudaFuncSetAttribute(k_recover_c<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = (int)std::min<uint64_t>((N + warps_per_cta - 1) / warps_per_cta, (uint64_t)ctx->sm_count * 4);
  k_recover_c<F><<<grid, 32 * warps_per_cta, smem, st>>>(d_shares, N, si, sj, t, dal.as<E>(), d_f, d_e, d_status,
                                                         ctx->d_count, d_pending, d_n_pending, quick);
  CKL();
  unsigned long long bad = 0;
  CK(cudaMemcpyAsync(&bad, ctx->d_count, sizeof(bad), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));  // also keeps `dal` alive until the kernel is done
  if (n_failed) *n_failed = bad;
  if (bad) return fail(ctx, SCLGPU_ECORRECT, "could not correct shares");
  return SCLGPU_OK;
}

template <class F>
static int recover_c_dev(sclgpu_ctx* ctx, const void* d_shares, uint64_t N, uint32_t n, int layout, const void* alphas,
                         void* d_f, void* d_e, uint8_t* d_status, uint64_t* n_failed) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (N && (!d_shares || !d_f || !d_e || !d_status)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  CK(cudaSetDevice(ctx->device));
  uint64_t si, sj;
  strides_for(layout, N, n, si, sj);
  return recover_c_on<F>(ctx, ctx->stream, (const E*)d_shares, N, n, si, sj, (const E*)alphas, (E*)d_f, (E*)d_e, d_status,
                         n_failed);
}

template <class F>
static int recover_c_host(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n, const void* alphas, void* f,
                          void* e, uint8_t* status, uint64_t* n_failed) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (N && (!shares || !f || !e || !status)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (n == 0) return fail(ctx, SCLGPU_EINVAL, "shamirRecoverC needs at least one share");
  CK(cudaSetDevice(ctx->device));
  const uint32_t t = (n - 1) / 3, np = 3 * t + 1;
  HostOp hop(ctx);
  void *dsh, *df, *de, *dst;
  RET(hop.up(shares, (size_t)N * n * sizeof(E), &dsh));
  RET(hop.dev((size_t)N * np * sizeof(E), &df));
  RET(hop.dev((size_t)N * (t + 1) * sizeof(E), &de));
  RET(hop.dev((size_t)N, &dst));
  const int rc = recover_c_on<F>(ctx, hop.st, (const E*)dsh, N, n, 1, n, (const E*)alphas, (E*)df, (E*)de, (uint8_t*)dst,
                                 n_failed);
  if (rc != SCLGPU_OK && rc != SCLGPU_ECORRECT) return rc;
  if (N) {
    CK(ctx->stager.d2h(hop.st, f, df, (size_t)N * np * sizeof(E)));
    CK(ctx->stager.d2h(hop.st, e, de, (size_t)N * (t + 1) * sizeof(E)));
    RET(hop.down(status, dst, (size_t)N));
  }
  return rc;
}
extern "C" int sclgpu_fp61_recover_c(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, const uint64_t* a, uint64_t* f, uint64_t* e, uint8_t* st, uint64_t* nf) { return guarded(c, [&] { return recover_c_host<F61>(c, s, N, n, a, f, e, st, nf); }); }
extern "C" int sclgpu_fp127_recover_c(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t n, const void* a, void* f, void* e, uint8_t* st, uint64_t* nf) { return guarded(c, [&] { return recover_c_host<F127>(c, s, N, n, a, f, e, st, nf); }); }
extern "C" int sclgpu_fp61_recover_c_dev(sclgpu_ctx* c, const uint64_t* s, uint64_t N, uint32_t n, int layout, const uint64_t* a, uint64_t* f, uint64_t* e, uint8_t* st, uint64_t* nf) { return guarded(c, [&] { return recover_c_dev<F61>(c, s, N, n, layout, a, f, e, st, nf); }); }
extern "C" int sclgpu_fp127_recover_c_dev(sclgpu_ctx* c, const void* s, uint64_t N, uint32_t n, int layout, const void* a, void* f, void* e, uint8_t* st, uint64_t* nf) { return guarded(c, [&] { return recover_c_dev<F127>(c, s, N, n, layout, a, f, e, st, nf); }); }

// ------------------------------------------------------------------ vector ops
template <class F, int OP>
static int binop_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* a, const typename F::E* b,
                    uint64_t n, typename F::E* out) {
  if (n == 0) return SCLGPU_OK;
  if constexpr (F::BYTES == 8) {
    const bool al = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) |
                      reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (al && n >= 2) {
      const uint64_t n2 = n >> 1;
      k_vec_binop61_v2<OP><<<grid_for(ctx, n2, 256, 8), 256, 0, st>>>(
          reinterpret_cast<const ulonglong2*>(a), reinterpret_cast<const ulonglong2*>(b), n2,
          reinterpret_cast<ulonglong2*>(out));
      CKL();
      if (n & 1) {
        k_vec_binop<F, OP><<<1, 32, 0, st>>>(a + n - 1, b + n - 1, 1, out + n - 1);
        CKL();
      }
      return SCLGPU_OK;
    }
  }
  k_vec_binop<F, OP><<<grid_for(ctx, n, 256, 8), 256, 0, st>>>(a, b, n, out);
  CKL();
  return SCLGPU_OK;
}

template <class F>
static int muladd_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* e, const typename F::E* b,
                     const typename F::E* d, const typename F::E* a, const typename F::E* c, uint64_t n,
                     typename F::E* z) {
  if (n == 0) return SCLGPU_OK;
  const uint64_t work = F::BYTES == 8 ? (n + 1) / 2 : n;
  k_vec_muladd<F><<<grid_for(ctx, work, 256, 8), 256, 0, st>>>(e, b, d, a, c, n, z);
  CKL();
  return SCLGPU_OK;
}

template <class F>
static int dot_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* a, const typename F::E* b,
                  uint64_t n, typename F::E* d_out) {
  typedef typename F::E E;
  const int grid = std::min(grid_for(ctx, n, 256, 8), kMaxPartials);
  k_dot_partial<F><<<grid, 256, 0, st>>>(a, b, n, reinterpret_cast<E*>(ctx->d_partial));
  CKL();
  k_sum_final<F><<<1, 256, 0, st>>>(reinterpret_cast<const E*>(ctx->d_partial), (uint32_t)grid, d_out);
  CKL();
  return SCLGPU_OK;
}

// op: 0 add 1 sub 2 mul 3 scale 4 dot 5 sum 6 muladd
template <class F>
static int vec_dev(sclgpu_ctx* ctx, int op, const void* a, const void* b, const void* d, const void* a2,
                   const void* c, uint64_t n, void* out) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (n && (!a || !out)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  cudaStream_t st = ctx->stream;
  switch (op) {
    case 0: return binop_on<F, 0>(ctx, st, (const E*)a, (const E*)b, n, (E*)out);
    case 1: return binop_on<F, 1>(ctx, st, (const E*)a, (const E*)b, n, (E*)out);
    case 2: return binop_on<F, 2>(ctx, st, (const E*)a, (const E*)b, n, (E*)out);
    case 3: {
      if (!b) return fail(ctx, SCLGPU_EINVAL, "null scalar");
      if (n == 0) return SCLGPU_OK;
      k_vec_scale<F><<<grid_for(ctx, n, 256, 8), 256, 0, st>>>((const E*)a, *(const E*)b, n, (E*)out);
      CKL();
      return SCLGPU_OK;
    }
    case 4: return dot_on<F>(ctx, st, (const E*)a, (const E*)b, n, (E*)out);
    case 5: return dot_on<F>(ctx, st, (const E*)a, nullptr, n, (E*)out);
    case 6: return muladd_on<F>(ctx, st, (const E*)a, (const E*)b, (const E*)d, (const E*)a2, (const E*)c, n, (E*)out);
    default: return fail(ctx, SCLGPU_EINVAL, "bad op");
  }
}

template <class F>
static int vec_host(sclgpu_ctx* ctx, int op, const void* a, const void* b, const void* d, const void* a2,
                    const void* c, uint64_t n, void* out) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (!out || (n && !a)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  HostOp hop(ctx);
  const size_t bytes = n * sizeof(E);
  void *da = nullptr, *db = nullptr, *dd = nullptr, *da2 = nullptr, *dc = nullptr, *dout = nullptr;
  RET(hop.up(a, bytes, &da));
  if (op == 0 || op == 1 || op == 2 || op == 4 || op == 6) {
    if (n && !b) return fail(ctx, SCLGPU_EINVAL, "null argument");
    RET(hop.up(b, bytes, &db));
  }
  if (op == 6) {
    if (n && (!d || !a2 || !c)) return fail(ctx, SCLGPU_EINVAL, "null argument");
    RET(hop.up(d, bytes, &dd));
    RET(hop.up(a2, bytes, &da2));
    RET(hop.up(c, bytes, &dc));
  }
  const size_t out_bytes = (op == 4 || op == 5) ? sizeof(E) : bytes;
  RET(hop.dev(out_bytes, &dout));
  cudaStream_t saved = ctx->stream;
  ctx->stream = hop.st;
  const int rc = vec_dev<F>(ctx, op, da, op == 3 ? b : db, dd, da2, dc, n, dout);
  ctx->stream = saved;
  RET(rc);
  return hop.down(out, dout, out_bytes);
}

#define SCLGPU_VEC_API(SUF, F, PTR, CPTR)                                                                                              \
  extern "C" int sclgpu_##SUF##_vec_add(sclgpu_ctx* c, CPTR a, CPTR b, uint64_t n, PTR o) { return guarded(c, [&] { return vec_host<F>(c, 0, a, b, 0, 0, 0, n, o); }); } \
  extern "C" int sclgpu_##SUF##_vec_sub(sclgpu_ctx* c, CPTR a, CPTR b, uint64_t n, PTR o) { return guarded(c, [&] { return vec_host<F>(c, 1, a, b, 0, 0, 0, n, o); }); } \
  extern "C" int sclgpu_##SUF##_vec_mul(sclgpu_ctx* c, CPTR a, CPTR b, uint64_t n, PTR o) { return guarded(c, [&] { return vec_host<F>(c, 2, a, b, 0, 0, 0, n, o); }); } \
  extern "C" int sclgpu_##SUF##_vec_scale(sclgpu_ctx* c, CPTR a, CPTR s, uint64_t n, PTR o) { return guarded(c, [&] { return vec_host<F>(c, 3, a, s, 0, 0, 0, n, o); }); } \
  extern "C" int sclgpu_##SUF##_vec_muladd(sclgpu_ctx* c, CPTR e, CPTR b, CPTR d, CPTR a, CPTR cc, uint64_t n, PTR z) { return guarded(c, [&] { return vec_host<F>(c, 6, e, b, d, a, cc, n, z); }); } \
  extern "C" int sclgpu_##SUF##_dot(sclgpu_ctx* c, CPTR a, CPTR b, uint64_t n, PTR o) { return guarded(c, [&] { return vec_host<F>(c, 4, a, b, 0, 0, 0, n, o); }); }     \
  extern "C" int sclgpu_##SUF##_sum(sclgpu_ctx* c, CPTR a, uint64_t n, PTR o) { return guarded(c, [&] { return vec_host<F>(c, 5, a, 0, 0, 0, 0, n, o); }); }             \
  extern "C" int sclgpu_##SUF##_vec_add_dev(sclgpu_ctx* c, CPTR a, CPTR b, uint64_t n, PTR o) { return guarded(c, [&] { return vec_dev<F>(c, 0, a, b, 0, 0, 0, n, o); }); } \
  extern "C" int sclgpu_##SUF##_vec_sub_dev(sclgpu_ctx* c, CPTR a, CPTR b, uint64_t n, PTR o) { return guarded(c, [&] { return vec_dev<F>(c, 1, a, b, 0, 0, 0, n, o); }); } \
  extern "C" int sclgpu_##SUF##_vec_mul_dev(sclgpu_ctx* c, CPTR a, CPTR b, uint64_t n, PTR o) { return guarded(c, [&] { return vec_dev<F>(c, 2, a, b, 0, 0, 0, n, o); }); } \
  extern "C" int sclgpu_##SUF##_vec_scale_dev(sclgpu_ctx* c, CPTR a, CPTR s, uint64_t n, PTR o) { return guarded(c, [&] { return vec_dev<F>(c, 3, a, s, 0, 0, 0, n, o); }); } \
  extern "C" int sclgpu_##SUF##_vec_muladd_dev(sclgpu_ctx* c, CPTR e, CPTR b, CPTR d, CPTR a, CPTR cc, uint64_t n, PTR z) { return guarded(c, [&] { return vec_dev<F>(c, 6, e, b, d, a, cc, n, z); }); } \
  extern "C" int sclgpu_##SUF##_dot_dev(sclgpu_ctx* c, CPTR a, CPTR b, uint64_t n, PTR o) { return guarded(c, [&] { return vec_dev<F>(c, 4, a, b, 0, 0, 0, n, o); }); }  \
  extern "C" int sclgpu_##SUF##_sum_dev(sclgpu_ctx* c, CPTR a, uint64_t n, PTR o) { return guarded(c, [&] { return vec_dev<F>(c, 5, a, 0, 0, 0, 0, n, o); }); }

// Vector::equals (vector.h:358-375): *equal = 1 iff all n elements agree (sizes are the caller's check)
template <class F>
static int vec_equal_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* a, const typename F::E* b, uint64_t n, int* equal) {
  *equal = 1;
  if (n == 0) return SCLGPU_OK;
  CK(cudaMemsetAsync(ctx->d_count, 0, sizeof(unsigned long long), st));
  k_vec_mismatch<F><<<grid_for(ctx, n, 256, 8), 256, 0, st>>>(a, b, n, ctx->d_count);
  CKL();
  unsigned long long bad = 0;
  CK(cudaMemcpyAsync(&bad, ctx->d_count, sizeof(bad), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *equal = bad == 0;
  return SCLGPU_OK;
}
template <class F>
static int vec_equal_dev(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, int* equal) {
  typedef typename F::E E;
  if (!ctx || !equal || (n && (!a || !b))) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return vec_equal_on<F>(ctx, ctx->stream, (const E*)a, (const E*)b, n, equal);
}
template <class F>
static int vec_equal_host(sclgpu_ctx* ctx, const void* a, const void* b, uint64_t n, int* equal) {
  typedef typename F::E E;
  if (!ctx || !equal || (n && (!a || !b))) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  HostOp hop(ctx);
  void *da, *db;
  RET(hop.up(a, n * sizeof(E), &da));
  RET(hop.up(b, n * sizeof(E), &db));
  return vec_equal_on<F>(ctx, hop.st, (const E*)da, (const E*)db, n, equal);
}
extern "C" int sclgpu_fp61_vec_equal(sclgpu_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t n, int* eq) { return guarded(c, [&] { return vec_equal_host<F61>(c, a, b, n, eq); }); }
extern "C" int sclgpu_fp127_vec_equal(sclgpu_ctx* c, const void* a, const void* b, uint64_t n, int* eq) { return guarded(c, [&] { return vec_equal_host<F127>(c, a, b, n, eq); }); }
extern "C" int sclgpu_fp61_vec_equal_dev(sclgpu_ctx* c, const uint64_t* a, const uint64_t* b, uint64_t n, int* eq) { return guarded(c, [&] { return vec_equal_dev<F61>(c, a, b, n, eq); }); }
extern "C" int sclgpu_fp127_vec_equal_dev(sclgpu_ctx* c, const void* a, const void* b, uint64_t n, int* eq) { return guarded(c, [&] { return vec_equal_dev<F127>(c, a, b, n, eq); }); }

SCLGPU_VEC_API(fp61, F61, uint64_t*, const uint64_t*)
SCLGPU_VEC_API(fp127, F127, void*, const void*)

// ------------------------------------------------------------------ matrix
template <class F>
static int matvec_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* A, uint32_t rows, uint32_t cols,
                     const typename F::E* x, typename F::E* y) {
  const int grid = (int)std::min<uint64_t>(rows, (uint64_t)ctx->sm_count * 8);
  if constexpr (F::BYTES == 8) {
    if ((cols & 1) == 0 && ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(x)) & 15) == 0) {
      if (cols % 512 == 0 && (uint64_t)rows * cols >= (1ull << 22) && !env_flag("SCLGPU_MATVEC_WARP")) {
        // long rows: chunked sweep in memory order + per-row finish (partials in stream-ordered scratch)
        const uint32_t cpr = cols / 512;
        const uint64_t n_chunks = (uint64_t)rows * cpr;
        StreamBuf part(ctx, st);
        CK(part.alloc(n_chunks * sizeof(uint64_t)));
        const int cgrid = (int)std::min<uint64_t>((n_chunks + 7) / 8, (uint64_t)ctx->sm_count * 8);
        k_matvec61_chunks<<<cgrid, 256, 0, st>>>(A, n_chunks, cpr, x, part.as<uint64_t>());
        CKL();
        k_matvec61_finish<<<(rows + 255) / 256, 256, 0, st>>>(part.as<uint64_t>(), rows, cpr, y);
      } else if (cols >= 256) {  // one warp per row
        const int wgrid = (int)std::min<uint64_t>(((uint64_t)rows + 7) / 8, (uint64_t)ctx->sm_count * 8);
        k_matvec61_warp<<<wgrid, 256, 0, st>>>(A, rows, cols, x, y);
      } else {
        k_matvec61_v2<<<grid, 256, 0, st>>>(A, rows, cols, x, y);
      }
      CKL();
      return SCLGPU_OK;
    }
  }
  k_matvec<F><<<grid, 256, 0, st>>>(A, rows, cols, x, y);
  CKL();
  return SCLGPU_OK;
}

template <class F>
static int matvec_dev(sclgpu_ctx* ctx, const void* A, uint32_t rows, uint32_t cols, const void* x, void* y) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (rows == 0 || cols == 0) return fail(ctx, SCLGPU_EINVAL, "n or m cannot be 0");
  if (!A || !x || !y) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return matvec_on<F>(ctx, ctx->stream, (const E*)A, rows, cols, (const E*)x, (E*)y);
}

template <class F>
static int matvec_host(sclgpu_ctx* ctx, const void* A, uint32_t rows, uint32_t cols, const void* x, void* y) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (rows == 0 || cols == 0) return fail(ctx, SCLGPU_EINVAL, "n or m cannot be 0");
  if (!A || !x || !y) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  HostOp hop(ctx);
  void *dA, *dx, *dy;
  RET(hop.up(A, (size_t)rows * cols * sizeof(E), &dA));
  RET(hop.up(x, (size_t)cols * sizeof(E), &dx));
  RET(hop.dev((size_t)rows * sizeof(E), &dy));
  RET(matvec_on<F>(ctx, hop.st, (const E*)dA, rows, cols, (const E*)dx, (E*)dy));
  return hop.down(y, dy, (size_t)rows * sizeof(E));
}
// Matrix::multiply(Matrix), matrix.h:476-495: C (rows x cols) = A (rows x inner) * B (inner x cols), row-major
template <class F>
static int matmul_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* A, uint32_t rows, uint32_t inner,
                     const typename F::E* B, uint32_t cols, typename F::E* C) {
  ctx->launches++;
  cudaError_t e;
  const uint64_t work = (uint64_t)rows * cols * inner;
  const bool big = work >= (F::BYTES == 8 ? (1ull << 18) : (1ull << 16)) && !env_flag("SCLGPU_MATMUL_GENERIC");
  const bool v1 = env_flag("SCLGPU_MATMUL_V1");  // the cp.async form (Fp61: even inner dimension, aligned A)
  if (big && !(v1 && F::BYTES == 8 && ((inner & 1) || (reinterpret_cast<uintptr_t>(A) & 15)))) {
    void* scratch = nullptr;
    size_t bytes;
    if constexpr (F::BYTES == 8) bytes = v1 ? matmul61_image_bytes(inner, cols) : matmul61_ws_scratch_bytes(rows, inner, cols);
    else bytes = v1 ? matmul127_image_bytes(inner, cols) : matmul127_ws_scratch_bytes(rows, inner, cols);
    CK(cudaMallocAsync(&scratch, bytes, st));
    ctx->launches += v1 ? 1 : 2;
    if constexpr (F::BYTES == 8) {
      e = v1 ? matmul61_tc_launch(st, ctx->sm_count, A, rows, inner, B, cols, (uint8_t*)scratch, C)
             : matmul61_ws_launch(st, A, rows, inner, B, cols, (uint8_t*)scratch, C);
    } else {
      e = v1 ? matmul127_tc_launch(st, ctx->sm_count, A, rows, inner, B, cols, (uint8_t*)scratch, C)
             : matmul127_ws_launch(st, A, rows, inner, B, cols, (uint8_t*)scratch, C);
    }
    cudaFreeAsync(scratch, st);
  } else if constexpr (F::BYTES == 8) {
    e = matmul61_generic_launch(st, ctx->sm_count, A, rows, inner, B, cols, C);
  } else {
    e = matmul127_generic_launch(st, ctx->sm_count, A, rows, inner, B, cols, C);
  }
  if (e != cudaSuccess) return cuda_fail(ctx, e, "launch");
  return SCLGPU_OK;
}

template <class F>
static int matmul_dev(sclgpu_ctx* ctx, const void* A, uint32_t rows, uint32_t inner, const void* B, uint32_t cols, void* C) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (rows == 0 || inner == 0 || cols == 0) return fail(ctx, SCLGPU_EINVAL, "n or m cannot be 0");
  if (!A || !B || !C) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return matmul_on<F>(ctx, ctx->stream, (const E*)A, rows, inner, (const E*)B, cols, (E*)C);
}

template <class F>
static int matmul_host(sclgpu_ctx* ctx, const void* A, uint32_t rows, uint32_t inner, const void* B, uint32_t cols, void* C) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (rows == 0 || inner == 0 || cols == 0) return fail(ctx, SCLGPU_EINVAL, "n or m cannot be 0");
  if (!A || !B || !C) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  HostOp hop(ctx);
  void *dA, *dB, *dC;
  RET(hop.up(A, (size_t)rows * inner * sizeof(E), &dA));
  RET(hop.up(B, (size_t)inner * cols * sizeof(E), &dB));
  RET(hop.dev((size_t)rows * cols * sizeof(E), &dC));
  RET(matmul_on<F>(ctx, hop.st, (const E*)dA, rows, inner, (const E*)dB, cols, (E*)dC));
  return hop.down(C, dC, (size_t)rows * cols * sizeof(E));
}
extern "C" int sclgpu_fp61_matmul(sclgpu_ctx* c, const uint64_t* A, uint32_t r, uint32_t k, const uint64_t* B, uint32_t n, uint64_t* C) { return guarded(c, [&] { return matmul_host<F61>(c, A, r, k, B, n, C); }); }
extern "C" int sclgpu_fp127_matmul(sclgpu_ctx* c, const void* A, uint32_t r, uint32_t k, const void* B, uint32_t n, void* C) { return guarded(c, [&] { return matmul_host<F127>(c, A, r, k, B, n, C); }); }
extern "C" int sclgpu_fp61_matmul_dev(sclgpu_ctx* c, const uint64_t* A, uint32_t r, uint32_t k, const uint64_t* B, uint32_t n, uint64_t* C) { return guarded(c, [&] { return matmul_dev<F61>(c, A, r, k, B, n, C); }); }
extern "C" int sclgpu_fp127_matmul_dev(sclgpu_ctx* c, const void* A, uint32_t r, uint32_t k, const void* B, uint32_t n, void* C) { return guarded(c, [&] { return matmul_dev<F127>(c, A, r, k, B, n, C); }); }

extern "C" int sclgpu_fp61_matvec(sclgpu_ctx* c, const uint64_t* A, uint32_t r, uint32_t k, const uint64_t* x, uint64_t* y) { return guarded(c, [&] { return matvec_host<F61>(c, A, r, k, x, y); }); }
extern "C" int sclgpu_fp127_matvec(sclgpu_ctx* c, const void* A, uint32_t r, uint32_t k, const void* x, void* y) { return guarded(c, [&] { return matvec_host<F127>(c, A, r, k, x, y); }); }
extern "C" int sclgpu_fp61_matvec_dev(sclgpu_ctx* c, const uint64_t* A, uint32_t r, uint32_t k, const uint64_t* x, uint64_t* y) { return guarded(c, [&] { return matvec_dev<F61>(c, A, r, k, x, y); }); }
extern "C" int sclgpu_fp127_matvec_dev(sclgpu_ctx* c, const void* A, uint32_t r, uint32_t k, const void* x, void* y) { return guarded(c, [&] { return matvec_dev<F127>(c, A, r, k, x, y); }); }

template <class F>
static int vandermonde_host(sclgpu_ctx* ctx, uint32_t n, uint32_t m, void* out) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (n == 0 || m == 0) return fail(ctx, SCLGPU_EINVAL, "n or m cannot be 0");
  if (!out) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  HostOp hop(ctx);
  void* dv;
  RET(hop.dev((size_t)n * m * sizeof(E), &dv));
  k_vandermonde<F><<<(n + 127) / 128, 128, 0, hop.st>>>(n, m, (E*)dv);
  CKL();
  return hop.down(out, dv, (size_t)n * m * sizeof(E));
}
extern "C" int sclgpu_fp61_vandermonde(sclgpu_ctx* c, uint32_t n, uint32_t m, uint64_t* o) { return guarded(c, [&] { return vandermonde_host<F61>(c, n, m, o); }); }
extern "C" int sclgpu_fp127_vandermonde(sclgpu_ctx* c, uint32_t n, uint32_t m, void* o) { return guarded(c, [&] { return vandermonde_host<F127>(c, n, m, o); }); }

// Matrix::vandermonde(n, m, xs) with the caller's nodes (matrix.h:445-460)
template <class F>
static int vandermonde_xs_host(sclgpu_ctx* ctx, uint32_t n, uint32_t m, const void* xs, uint32_t n_xs, void* out) {
  typedef typename F::E E;
  if (!ctx) return SCLGPU_EINVAL;
  if (n_xs != n) return fail(ctx, SCLGPU_EINVAL, "|xs| != number of rows");
  if (n == 0 || m == 0) return SCLGPU_OK;  // the reference builds an empty matrix here (no "n or m cannot be 0" check)
  if (!xs || !out) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  HostOp hop(ctx);
  void *dx, *dv;
  RET(hop.up(xs, (size_t)n * sizeof(E), &dx));
  RET(hop.dev((size_t)n * m * sizeof(E), &dv));
  k_vandermonde_xs<F><<<(n + 127) / 128, 128, 0, hop.st>>>(n, m, (const E*)dx, (E*)dv);
  CKL();
  return hop.down(out, dv, (size_t)n * m * sizeof(E));
}
extern "C" int sclgpu_fp61_vandermonde_xs(sclgpu_ctx* c, uint32_t n, uint32_t m, const uint64_t* xs, uint32_t nx, uint64_t* o) { return guarded(c, [&] { return vandermonde_xs_host<F61>(c, n, m, xs, nx, o); }); }
extern "C" int sclgpu_fp127_vandermonde_xs(sclgpu_ctx* c, uint32_t n, uint32_t m, const void* xs, uint32_t nx, void* o) { return guarded(c, [&] { return vandermonde_xs_host<F127>(c, n, m, xs, nx, o); }); }

// Polynomial::evaluate (poly.h:56-64) of N polynomials at n caller-chosen points.  xs: HOST pointer (n points).
template <class F>
static int poly_eval_on(sclgpu_ctx* ctx, cudaStream_t st, const typename F::E* d_coeffs, uint64_t N, uint32_t t,
                        const typename F::E* d_xs, uint32_t n, typename F::E* d_out, uint64_t si, uint64_t sj) {
  typedef typename F::E E;
  if (N == 0 || n == 0) return SCLGPU_OK;
  const size_t smem = (size_t)n * sizeof(E);
  if (smem > 200 * 1024) return fail(ctx, SCLGPU_EINVAL, "poly_evaluate: more than 200 KiB of evaluation points");
  if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k_poly_eval<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_poly_eval<F><<<grid_for(ctx, N, 256, 8), 256, smem, st>>>(d_coeffs, N, t, d_xs, n, d_out, si, sj);
  CKL();
  return SCLGPU_OK;
}
template <class F>
static int poly_eval_dev(sclgpu_ctx* ctx, const void* d_coeffs, uint64_t N, uint32_t t, const void* xs, uint32_t n,
                         void* d_out, int layout) {
  typedef typename F::E E;
  if (!ctx || ((!d_coeffs || !d_out || !xs) && N && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  if (layout != SCLGPU_PARTY_MAJOR && layout != SCLGPU_SECRET_MAJOR) return fail(ctx, SCLGPU_EINVAL, "bad layout");
  CK(cudaSetDevice(ctx->device));
  if (N == 0 || n == 0) return SCLGPU_OK;
  StreamBuf dx(ctx, ctx->stream);
  CK(dx.alloc((size_t)n * sizeof(E)));
  CK(cudaMemcpyAsync(dx.p, xs, (size_t)n * sizeof(E), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));  // xs is the caller's host buffer
  uint64_t si, sj;
  strides_for(layout, N, n, si, sj);
  return poly_eval_on<F>(ctx, ctx->stream, (const E*)d_coeffs, N, t, dx.as<E>(), n, (E*)d_out, si, sj);
}
// host form: coeffs [N][t+1] (row j = the coefficients of polynomial j, constant term first, as
// Polynomial::coefficients() holds them), out [N][n]
template <class F>
static int poly_eval_host(sclgpu_ctx* ctx, const void* coeffs, uint64_t N, uint32_t t, const void* xs, uint32_t n, void* out) {
  typedef typename F::E E;
  if (!ctx || ((!coeffs || !out || !xs) && N && n)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  if (N == 0 || n == 0) return SCLGPU_OK;
  const uint64_t m = (uint64_t)t + 1;
  uint64_t chunk = std::max<uint64_t>((256ull << 20) / (std::max<uint64_t>(m, n) * sizeof(E)), 1024);
  chunk = std::min(chunk, std::min(N, kHostChunk));
  const int nbuf = N > chunk ? 2 : 1;
  PoolScope pool_scope(ctx);
  PoolBuf din[2], dpl[2], dpm[2], dsm[2], dx;
  CK(dx.alloc((size_t)n * sizeof(E)));
  for (int k = 0; k < nbuf; ++k) {
    CK(din[k].alloc(chunk * m * sizeof(E)));
    CK(dpl[k].alloc(chunk * m * sizeof(E)));
    CK(dpm[k].alloc(chunk * n * sizeof(E)));
    CK(dsm[k].alloc(chunk * n * sizeof(E)));
  }
  CK(ctx->stager.h2d(ctx->pipe[0], dx.p, xs, (size_t)n * sizeof(E)));
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(ctx->stager.drain());
  const E* hc = reinterpret_cast<const E*>(coeffs);
  E* ho = reinterpret_cast<E*>(out);
  int k = 0;
  for (uint64_t c0 = 0; c0 < N; c0 += chunk, k ^= (nbuf - 1)) {
    const uint64_t nc = std::min(chunk, N - c0);
    cudaStream_t st = ctx->pipe[k];
    CK(ctx->stager.h2d(st, din[k].p, hc + c0 * m, nc * m * sizeof(E)));
    RET(transpose_on<E>(ctx, st, din[k].as<E>(), nc, m, dpl[k].as<E>()));          // [nc][t+1] -> planes [t+1][nc]
    RET(poly_eval_on<F>(ctx, st, dpl[k].as<E>(), nc, t, dx.as<E>(), n, dpm[k].as<E>(), nc, 1));
    RET(transpose_on<E>(ctx, st, dpm[k].as<E>(), n, nc, dsm[k].as<E>()));          // [n][nc] -> [nc][n]
    CK(ctx->stager.d2h(st, ho + c0 * n, dsm[k].p, nc * n * sizeof(E)));
  }
  CK(cudaStreamSynchronize(ctx->pipe[0]));
  CK(cudaStreamSynchronize(ctx->pipe[1]));
  CK(ctx->stager.drain());
  return SCLGPU_OK;
}
extern "C" int sclgpu_fp61_poly_evaluate(sclgpu_ctx* c, const uint64_t* k, uint64_t N, uint32_t t, const uint64_t* xs, uint32_t n, uint64_t* o) { return guarded(c, [&] { return poly_eval_host<F61>(c, k, N, t, xs, n, o); }); }
extern "C" int sclgpu_fp127_poly_evaluate(sclgpu_ctx* c, const void* k, uint64_t N, uint32_t t, const void* xs, uint32_t n, void* o) { return guarded(c, [&] { return poly_eval_host<F127>(c, k, N, t, xs, n, o); }); }
extern "C" int sclgpu_fp61_poly_evaluate_dev(sclgpu_ctx* c, const uint64_t* k, uint64_t N, uint32_t t, const uint64_t* xs, uint32_t n, uint64_t* o, int layout) { return guarded(c, [&] { return poly_eval_dev<F61>(c, k, N, t, xs, n, o, layout); }); }
extern "C" int sclgpu_fp127_poly_evaluate_dev(sclgpu_ctx* c, const void* k, uint64_t N, uint32_t t, const void* xs, uint32_t n, void* o, int layout) { return guarded(c, [&] { return poly_eval_dev<F127>(c, k, N, t, xs, n, o, layout); }); }

// Matrix::transpose (matrix.h:344-355) on a host matrix
template <class E>
static int transpose_host(sclgpu_ctx* ctx, const void* in, uint64_t rows, uint64_t cols, void* out) {
  if (!ctx || ((!in || !out) && rows && cols)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  if (rows == 0 || cols == 0) return SCLGPU_OK;
  HostOp hop(ctx);
  void *di, *dout;
  RET(hop.up(in, rows * cols * sizeof(E), &di));
  RET(hop.dev(rows * cols * sizeof(E), &dout));
  RET(transpose_on<E>(ctx, hop.st, (const E*)di, rows, cols, (E*)dout));
  return hop.down(out, dout, rows * cols * sizeof(E));
}
extern "C" int sclgpu_fp61_transpose(sclgpu_ctx* c, const uint64_t* in, uint64_t rows, uint64_t cols, uint64_t* out) { return guarded(c, [&] { return transpose_host<uint64_t>(c, in, rows, cols, out); }); }
extern "C" int sclgpu_fp127_transpose(sclgpu_ctx* c, const void* in, uint64_t rows, uint64_t cols, void* out) { return guarded(c, [&] { return transpose_host<E127>(c, in, rows, cols, out); }); }

static int sclgpu_fp61_transpose_dev_impl(sclgpu_ctx* ctx, const uint64_t* in, uint64_t rows, uint64_t cols, uint64_t* out) {
  if (!ctx || ((!in || !out) && rows && cols)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return transpose_on<uint64_t>(ctx, ctx->stream, in, rows, cols, out);
}
static int sclgpu_fp127_transpose_dev_impl(sclgpu_ctx* ctx, const void* in, uint64_t rows, uint64_t cols, void* out) {
  if (!ctx || ((!in || !out) && rows && cols)) return fail(ctx, SCLGPU_EINVAL, "null argument");
  CK(cudaSetDevice(ctx->device));
  return transpose_on<E127>(ctx, ctx->stream, (const E127*)in, rows, cols, (E127*)out);
}

// ------------------------------------------------------------------ microbench
static int sclgpu_pipe_microbench_impl(sclgpu_ctx* ctx, int kind, uint32_t iters, double* ops_per_s) {
  if (!ctx || !ops_per_s || kind < 0 || kind > 4) return fail(ctx, SCLGPU_EINVAL, "bad argument");
  CK(cudaSetDevice(ctx->device));
  DevBuf sink;
  CK(sink.alloc(16));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  const int grid = ctx->sm_count, threads = 512;
  auto launch = [&](uint32_t it) {
    switch (kind) {
      case 0: k_pipe_bench<0><<<grid, threads, 0, ctx->stream>>>(it, sink.as<uint32_t>()); break;
      case 1: k_pipe_bench<1><<<grid, threads, 0, ctx->stream>>>(it, sink.as<uint32_t>()); break;
      case 2: k_pipe_bench<2><<<grid, threads, 0, ctx->stream>>>(it, sink.as<uint32_t>()); break;
      case 3: k_pipe_bench<3><<<grid, threads, 0, ctx->stream>>>(it, sink.as<uint32_t>()); break;
      default: k_pipe_bench<4><<<grid, threads, 0, ctx->stream>>>(it, sink.as<uint32_t>()); break;
    }
    ctx->launches++;
  };
  launch(iters / 8 + 1);  // warm-up
  CK(cudaEventRecord(e0, ctx->stream));
  launch(iters);
  CK(cudaEventRecord(e1, ctx->stream));
  CK(cudaEventSynchronize(e1));
  CK(cudaGetLastError());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double ops = (double)grid * threads * (double)iters * 64.0;  // 8 chains x 8 unroll
  *ops_per_s = ops / ((double)ms * 1e-3);
  return SCLGPU_OK;
}

// ------------------------------------------------------------------ guarded wrappers of the entry points above
// (every extern "C" function goes through guarded(): no C++ exception crosses the ABI)
extern "C" int sclgpu_prg_expand_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n_bytes, uint8_t* d_out) { return guarded(ctx, [&] { return sclgpu_prg_expand_dev_impl(ctx, seed, first_block, n_bytes, d_out); }); }
extern "C" int sclgpu_prg_expand(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n_bytes, uint8_t* out) { return guarded(ctx, [&] { return sclgpu_prg_expand_impl(ctx, seed, first_block, n_bytes, out); }); }
extern "C" int sclgpu_fp61_from_bytes_dev(sclgpu_ctx* ctx, const uint8_t* b, uint64_t n, uint64_t* o) { return guarded(ctx, [&] { return sclgpu_fp61_from_bytes_dev_impl(ctx, b, n, o); }); }
extern "C" int sclgpu_fp127_from_bytes_dev(sclgpu_ctx* ctx, const uint8_t* b, uint64_t n, void* o) { return guarded(ctx, [&] { return sclgpu_fp127_from_bytes_dev_impl(ctx, b, n, o); }); }
extern "C" int sclgpu_fp61_transpose_dev(sclgpu_ctx* ctx, const uint64_t* in, uint64_t rows, uint64_t cols, uint64_t* out) { return guarded(ctx, [&] { return sclgpu_fp61_transpose_dev_impl(ctx, in, rows, cols, out); }); }
extern "C" int sclgpu_fp127_transpose_dev(sclgpu_ctx* ctx, const void* in, uint64_t rows, uint64_t cols, void* out) { return guarded(ctx, [&] { return sclgpu_fp127_transpose_dev_impl(ctx, in, rows, cols, out); }); }
extern "C" int sclgpu_pipe_microbench(sclgpu_ctx* ctx, int kind, uint32_t iters, double* ops_per_s) { return guarded(ctx, [&] { return sclgpu_pipe_microbench_impl(ctx, kind, iters, ops_per_s); }); }
