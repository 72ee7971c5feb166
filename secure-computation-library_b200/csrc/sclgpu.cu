// sclgpu.cu -- host side of libsclgpu.so: context, launch logic, the C ABI of
// include/sclgpu.h.  Pure CUDA runtime; no torch types cross this boundary.
//
// One translation unit, laid out by subsystem: this file holds the context, the error / environment / scratch
// plumbing and the memory entry points; the rest is #included below in dependency order --
//   sclgpu_ops.inc      device-level operations (kernel dispatch) shared by the entry points
//   sclgpu_prg.inc      PRG, Vector::random / FF::random, FF::read
//   sclgpu_share.inc    shamirSecretShare (plain, coefficient planes, arrays, packets), additive sharing
//   sclgpu_recover.inc  Lagrange, shamirRecoverP / D / C, the single-launch step, the peer-memory gather
//   sclgpu_linalg.inc   Vector / Matrix operations, vandermonde, Polynomial::evaluate, microbench
// multi.cu (several GPUs behind one handle, asynchronous calls) is a separate unit on top of the public ABI.
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
static constexpr int kMaxPipes = 4;
struct sclgpu_ctx {
  int device = 0;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  cudaStream_t stream = nullptr;      // user stream for _dev entry points
  cudaStream_t pipe[kMaxPipes] = {};  // host-pointer pipelines (share / recoverP use up to kMaxPipes, the others two)
  cudaEvent_t pipe_ev[kMaxPipes] = {};
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
  // the cache holds what the ENVIRONMENT says (set or not), never a caller's default: two call sites may ask for the
  // same variable with different defaults (SCLGPU_SR_WARPS: same-batch and pipelined mode)
  static std::mutex mu;
  static std::map<std::string, std::pair<bool, int>> seen;
  std::lock_guard<std::mutex> lock(mu);
  auto it = seen.find(name);
  if (it == seen.end()) {
    const char* e = getenv(name);
    it = seen.emplace(name, std::make_pair(e != nullptr, e ? atoi(e) : 0)).first;
  }
  return it->second.first ? it->second.second : dflt;
}

// Depth and chunk size of the share / recoverP host pipelines.  A chunk's chain is small copy -> kernels -> big copy
// (share: secrets up, shares down; recoverP: shares up, secrets down).  When a share and a reconstruction run at the
// same time (sclgpu_*_async), the SMALL copy of one pipeline queues on the copy engine behind the BIG copies of the
// other.  Measured on B200 (2^25 secrets, n = 32, duplex, GB/s each way): 2 x 256 MiB 46.7, 3 x 128 43.9, 4 x 128 43.7,
// 4 x 64 41.1, 4 x 32 39.1 -- fewer, larger chunks win (a fixed cost of about 0.3 ms per chunk and direction), so the
// defaults stay at two chunks of 256 MiB in flight; the knobs remain for other hosts (DESIGN.md 8a').
static int host_pipes() {
  return std::min(std::max(env_int("SCLGPU_HOST_PIPES", 2), 1), kMaxPipes);
}
static uint64_t host_chunk_bytes() {
  return (uint64_t)std::min(std::max(env_int("SCLGPU_HOST_CHUNK_MB", 256), 1), 1024) << 20;
}
static cudaError_t sync_pipes(sclgpu_ctx* ctx) {
  cudaError_t first = cudaSuccess;
  for (int i = 0; i < kMaxPipes; ++i) {
    const cudaError_t e = cudaStreamSynchronize(ctx->pipe[i]);
    if (first == cudaSuccess) first = e;
  }
  return first;
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
  for (int i = 0; i < kMaxPipes && ok; ++i) {
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
  for (int i = 0; i < kMaxPipes; ++i) {
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
      for (int i = 0; i < kMaxPipes; ++i) cudaStreamSynchronize(g_pool_ctx->pipe[i]);
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

#include "sclgpu_ops.inc"
#include "sclgpu_prg.inc"
#include "sclgpu_share.inc"
#include "sclgpu_recover.inc"
#include "sclgpu_linalg.inc"

// ------------------------------------------------------------------ guarded wrappers of the entry points above
// (every extern "C" function goes through guarded(): no C++ exception crosses the ABI)
extern "C" int sclgpu_prg_expand_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n_bytes, uint8_t* d_out) { return guarded(ctx, [&] { return sclgpu_prg_expand_dev_impl(ctx, seed, first_block, n_bytes, d_out); }); }
extern "C" int sclgpu_prg_expand_bitsliced_dev(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n_bytes, uint8_t* d_out) { return guarded(ctx, [&] { return sclgpu_prg_expand_bitsliced_dev_impl(ctx, seed, first_block, n_bytes, d_out); }); }
extern "C" int sclgpu_prg_expand(sclgpu_ctx* ctx, const uint8_t seed[16], uint64_t first_block, uint64_t n_bytes, uint8_t* out) { return guarded(ctx, [&] { return sclgpu_prg_expand_impl(ctx, seed, first_block, n_bytes, out); }); }
extern "C" int sclgpu_fp61_from_bytes_dev(sclgpu_ctx* ctx, const uint8_t* b, uint64_t n, uint64_t* o) { return guarded(ctx, [&] { return sclgpu_fp61_from_bytes_dev_impl(ctx, b, n, o); }); }
extern "C" int sclgpu_fp127_from_bytes_dev(sclgpu_ctx* ctx, const uint8_t* b, uint64_t n, void* o) { return guarded(ctx, [&] { return sclgpu_fp127_from_bytes_dev_impl(ctx, b, n, o); }); }
extern "C" int sclgpu_fp61_transpose_dev(sclgpu_ctx* ctx, const uint64_t* in, uint64_t rows, uint64_t cols, uint64_t* out) { return guarded(ctx, [&] { return sclgpu_fp61_transpose_dev_impl(ctx, in, rows, cols, out); }); }
extern "C" int sclgpu_fp127_transpose_dev(sclgpu_ctx* ctx, const void* in, uint64_t rows, uint64_t cols, void* out) { return guarded(ctx, [&] { return sclgpu_fp127_transpose_dev_impl(ctx, in, rows, cols, out); }); }
extern "C" int sclgpu_pipe_microbench(sclgpu_ctx* ctx, int kind, uint32_t iters, double* ops_per_s) { return guarded(ctx, [&] { return sclgpu_pipe_microbench_impl(ctx, kind, iters, ops_per_s); }); }
