// multi.cu -- two host-side layers ABOVE the single-device C ABI (they only call the public entry points of
// include/sclgpu.h, so everything they do is also reachable by a caller's own threads):
//
//  * sclgpu_mctx: one handle over several GPUs of a box for an SCL-style caller (SCL is a single-process,
//    single-threaded library: include/scl/coro/runtime.h:126-163).  A batch entry point cuts [0, N) into contiguous
//    slices, one per device; slice g starts its PRG at first_block + lo_g * B (B = blocks per shamirSecretShare call,
//    shamir.h:56 + prg.cc:129-133) with the same seed, so the N results are exactly what N calls on ONE scl::util::PRG
//    return.  One worker thread per device runs the single-device host pipeline on its slice; the slices
//    of the caller's buffers are disjoint, no collective is involved.
//
//  * *_async: the host entry points are synchronous (they return when the last device-to-host copy has landed), so a
//    caller that shares batch k and reconstructs batch k-1 would use PCIe in one direction at a time.  An async call
//    runs the same pipeline on a companion context (own streams, own scratch) from a worker thread and returns at
//    once; sclgpu_wait / sclgpu_sync completes it.  share_async(k) + recover_p(k-1) then move data in both directions.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/sclgpu.h"

namespace {

uint64_t share_blocks(uint32_t element_bytes, uint32_t t) { return ((uint64_t)(t + 1) * element_bytes + 15) / 16; }

// slice g of [0, N) over `parts` devices, boundaries on multiples of `align`
void slice_of(uint64_t N, int parts, int g, uint64_t align, uint64_t& lo, uint64_t& hi) {
  const uint64_t groups = (N + align - 1) / align;
  lo = std::min(N, groups * (uint64_t)g / (uint64_t)parts * align);
  hi = std::min(N, groups * (uint64_t)(g + 1) / (uint64_t)parts * align);
}

}  // namespace

struct sclgpu_mctx {
  std::vector<sclgpu_ctx*> ctx;
  std::vector<int> device;
  std::string last_error;
};

namespace {

// run fn(g, lo, hi) for every device on its own thread; first failure wins
int for_each_slice(sclgpu_mctx* m, uint64_t N, uint64_t align, const std::function<int(int, uint64_t, uint64_t)>& fn) {
  const int G = (int)m->ctx.size();
  std::vector<int> rc(G, SCLGPU_OK);
  std::vector<std::thread> th;
  th.reserve(G);
  for (int g = 0; g < G; ++g) {
    uint64_t lo, hi;
    slice_of(N, G, g, align, lo, hi);
    if (lo == hi) continue;
    try {
      th.emplace_back([&, g, lo, hi] {
        try {
          rc[g] = fn(g, lo, hi);
        } catch (...) {
          rc[g] = SCLGPU_ECUDA;
        }
      });
    } catch (...) {  // thread creation failed: run the slice here
      rc[g] = fn(g, lo, hi);
    }
  }
  for (auto& t : th) t.join();
  for (int g = 0; g < G; ++g) {
    if (rc[g] != SCLGPU_OK && rc[g] != SCLGPU_EDETECT) {
      m->last_error = std::string("device ") + std::to_string(m->device[g]) + ": " + sclgpu_last_error(m->ctx[g]);
      return rc[g];
    }
  }
  for (int g = 0; g < G; ++g)
    if (rc[g] == SCLGPU_EDETECT) {
      m->last_error = sclgpu_last_error(m->ctx[g]);
      return SCLGPU_EDETECT;
    }
  return SCLGPU_OK;
}

int mfail(sclgpu_mctx* m, int code, const char* msg) {
  if (m) m->last_error = msg;
  return code;
}

}  // namespace

// the slicing rule itself (no device needed): slice `index` of `parts` over [0, n_units), boundaries on multiples of `align`
extern "C" int sclgpu_multi_slice(uint64_t n_units, int parts, int index, uint64_t align, uint64_t* lo, uint64_t* hi) {
  if (parts < 1 || index < 0 || index >= parts || align < 1 || !lo || !hi) return SCLGPU_EINVAL;
  slice_of(n_units, parts, index, align, *lo, *hi);
  return SCLGPU_OK;
}

extern "C" int sclgpu_multi_init(const int* devices, int n_devices, sclgpu_mctx** out) {
  if (!out) return SCLGPU_EINVAL;
  *out = nullptr;
  if (n_devices < 1) return SCLGPU_EINVAL;
  sclgpu_mctx* m = new (std::nothrow) sclgpu_mctx();
  if (!m) return SCLGPU_ENOMEM;
  try {
    for (int g = 0; g < n_devices; ++g) {
      const int dev = devices ? devices[g] : g;
      sclgpu_ctx* c = nullptr;
      const int rc = sclgpu_init(dev, &c);
      if (rc != SCLGPU_OK) {
        for (auto* p : m->ctx) sclgpu_destroy(p);
        delete m;
        return rc;  // no CPU fallback, no partial device set
      }
      m->ctx.push_back(c);
      m->device.push_back(dev);
    }
  } catch (...) {
    for (auto* p : m->ctx) sclgpu_destroy(p);
    delete m;
    return SCLGPU_ENOMEM;
  }
  *out = m;
  return SCLGPU_OK;
}

extern "C" void sclgpu_multi_destroy(sclgpu_mctx* m) {
  if (!m) return;
  for (auto* p : m->ctx) sclgpu_destroy(p);
  delete m;
}

extern "C" int sclgpu_multi_device_count(const sclgpu_mctx* m) { return m ? (int)m->ctx.size() : 0; }
extern "C" sclgpu_ctx* sclgpu_multi_context(sclgpu_mctx* m, int index) {
  return (m && index >= 0 && index < (int)m->ctx.size()) ? m->ctx[index] : nullptr;
}
extern "C" const char* sclgpu_multi_last_error(const sclgpu_mctx* m) { return m ? m->last_error.c_str() : "no context"; }

// ---- shamirSecretShare (shamir.h:52-68) x N over the devices; SCL's [N][n] layout
extern "C" int sclgpu_multi_fp61_shamir_share(sclgpu_mctx* m, const uint64_t* secrets, uint64_t N, uint32_t t, uint32_t n,
                                               const uint8_t seed[16], uint64_t first_block, uint64_t* shares) {
  if (!m || !seed || ((!secrets || !shares) && N && n)) return mfail(m, SCLGPU_EINVAL, "null argument");
  const uint64_t B = share_blocks(8, t);
  return for_each_slice(m, N, 2, [&](int g, uint64_t lo, uint64_t hi) {
    return sclgpu_fp61_shamir_share(m->ctx[g], secrets + lo, hi - lo, t, n, seed, first_block + lo * B, shares + lo * n);
  });
}
extern "C" int sclgpu_multi_fp127_shamir_share(sclgpu_mctx* m, const void* secrets, uint64_t N, uint32_t t, uint32_t n,
                                                const uint8_t seed[16], uint64_t first_block, void* shares) {
  if (!m || !seed || ((!secrets || !shares) && N && n)) return mfail(m, SCLGPU_EINVAL, "null argument");
  const uint64_t B = share_blocks(16, t);
  const uint8_t* s = static_cast<const uint8_t*>(secrets);
  uint8_t* o = static_cast<uint8_t*>(shares);
  return for_each_slice(m, N, 1, [&](int g, uint64_t lo, uint64_t hi) {
    return sclgpu_fp127_shamir_share(m->ctx[g], s + lo * 16, hi - lo, t, n, seed, first_block + lo * B, o + lo * n * 16);
  });
}

// ---- shamirRecoverP (shamir.h:82-104) x N
extern "C" int sclgpu_multi_fp61_recover_p(sclgpu_mctx* m, const uint64_t* shares, uint64_t N, uint32_t n, const uint64_t* alphas,
                                            const uint64_t* x, uint64_t* out) {
  if (!m || ((!shares && n) || !out) && N) return mfail(m, SCLGPU_EINVAL, "null argument");
  return for_each_slice(m, N, 2, [&](int g, uint64_t lo, uint64_t hi) {
    return sclgpu_fp61_recover_p(m->ctx[g], shares + lo * n, hi - lo, n, alphas, x, out + lo);
  });
}
extern "C" int sclgpu_multi_fp127_recover_p(sclgpu_mctx* m, const void* shares, uint64_t N, uint32_t n, const void* alphas,
                                             const void* x, void* out) {
  if (!m || ((!shares && n) || !out) && N) return mfail(m, SCLGPU_EINVAL, "null argument");
  const uint8_t* s = static_cast<const uint8_t*>(shares);
  uint8_t* o = static_cast<uint8_t*>(out);
  return for_each_slice(m, N, 1, [&](int g, uint64_t lo, uint64_t hi) {
    return sclgpu_fp127_recover_p(m->ctx[g], s + lo * n * 16, hi - lo, n, alphas, x, o + lo * 16);
  });
}

// ---- shamirRecoverD (shamir.h:117-155) x N: per-secret flags, the count summed over the devices
template <class Call>
static int multi_recover_d(sclgpu_mctx* m, uint64_t N, uint64_t* n_detected, Call call) {
  const int G = (int)m->ctx.size();
  std::vector<uint64_t> cnt(G, 0);
  const int rc = for_each_slice(m, N, 1, [&](int g, uint64_t lo, uint64_t hi) { return call(g, lo, hi, &cnt[g]); });
  uint64_t total = 0;
  for (uint64_t c : cnt) total += c;
  if (n_detected) *n_detected = total;
  return rc;
}
extern "C" int sclgpu_multi_fp61_recover_d(sclgpu_mctx* m, const uint64_t* shares, uint64_t N, uint32_t n_given, uint32_t t,
                                            const uint64_t* alphas, uint32_t n_alphas, uint32_t d, const uint64_t* x,
                                            uint64_t* out, uint8_t* err, uint64_t* n_detected) {
  if (!m || ((!shares && n_given) || !out || !err) && N) return mfail(m, SCLGPU_EINVAL, "null argument");
  return multi_recover_d(m, N, n_detected, [&](int g, uint64_t lo, uint64_t hi, uint64_t* c) {
    return sclgpu_fp61_recover_d(m->ctx[g], shares + lo * n_given, hi - lo, n_given, t, alphas, n_alphas, d, x, out + lo,
                                 err + lo, c);
  });
}
extern "C" int sclgpu_multi_fp127_recover_d(sclgpu_mctx* m, const void* shares, uint64_t N, uint32_t n_given, uint32_t t,
                                             const void* alphas, uint32_t n_alphas, uint32_t d, const void* x, void* out,
                                             uint8_t* err, uint64_t* n_detected) {
  if (!m || ((!shares && n_given) || !out || !err) && N) return mfail(m, SCLGPU_EINVAL, "null argument");
  const uint8_t* s = static_cast<const uint8_t*>(shares);
  uint8_t* o = static_cast<uint8_t*>(out);
  return multi_recover_d(m, N, n_detected, [&](int g, uint64_t lo, uint64_t hi, uint64_t* c) {
    return sclgpu_fp127_recover_d(m->ctx[g], s + lo * n_given * 16, hi - lo, n_given, t, alphas, n_alphas, d, x, o + lo * 16,
                                  err + lo, c);
  });
}

// ---- Vector::random (vector.h:508-519) over the devices: Fp61 draws two elements per block, so slices start on
// even elements
extern "C" int sclgpu_multi_fp61_random(sclgpu_mctx* m, const uint8_t seed[16], uint64_t first_block, uint64_t n, uint64_t* out) {
  if (!m || !seed || (!out && n)) return mfail(m, SCLGPU_EINVAL, "null argument");
  return for_each_slice(m, n, 2, [&](int g, uint64_t lo, uint64_t hi) {
    return sclgpu_fp61_random(m->ctx[g], seed, first_block + lo / 2, hi - lo, out + lo);
  });
}

// =================================================================== asynchronous host calls
namespace {

struct AsyncState {
  sclgpu_ctx* companion = nullptr;
  std::thread worker;
  bool pending = false;
  int rc = SCLGPU_OK;
  std::string error;
};
std::mutex g_async_mu;
std::map<sclgpu_ctx*, AsyncState*> g_async;

AsyncState* async_state(sclgpu_ctx* ctx, bool create) {
  std::lock_guard<std::mutex> lock(g_async_mu);
  auto it = g_async.find(ctx);
  if (it != g_async.end()) return it->second;
  if (!create) return nullptr;
  AsyncState* st = new (std::nothrow) AsyncState();
  if (!st) return nullptr;
  g_async[ctx] = st;
  return st;
}

int async_wait(AsyncState* st) {
  if (!st || !st->pending) return SCLGPU_OK;
  if (st->worker.joinable()) st->worker.join();
  st->pending = false;
  return st->rc;
}

// run `call(companion)` on the companion context of `ctx` from a worker thread
int async_launch(sclgpu_ctx* ctx, const std::function<int(sclgpu_ctx*)>& call) {
  if (!ctx) return SCLGPU_EINVAL;
  AsyncState* st = async_state(ctx, true);
  if (!st) return SCLGPU_ENOMEM;
  const int prev = async_wait(st);  // one asynchronous call in flight per context
  if (prev != SCLGPU_OK && prev != SCLGPU_EDETECT) return prev;
  if (!st->companion) {
    int dev = 0;
    sclgpu_device_index(ctx, &dev);
    const int rc = sclgpu_init(dev, &st->companion);
    if (rc != SCLGPU_OK) return rc;
  }
  st->rc = SCLGPU_OK;
  st->pending = true;
  try {
    st->worker = std::thread([st, call] {
      try {
        st->rc = call(st->companion);
        if (st->rc != SCLGPU_OK) st->error = sclgpu_last_error(st->companion);
      } catch (...) {
        st->rc = SCLGPU_ECUDA;
        st->error = "internal error in asynchronous call";
      }
    });
  } catch (...) {  // no thread: degrade to a synchronous call
    st->rc = call(st->companion);
    if (st->rc != SCLGPU_OK) st->error = sclgpu_last_error(st->companion);
  }
  return SCLGPU_OK;
}

}  // namespace

// called by sclgpu_sync / sclgpu_destroy (sclgpu.cu)
int sclgpu_async_complete(sclgpu_ctx* ctx, const char** error) {
  AsyncState* st = async_state(ctx, false);
  if (!st) return SCLGPU_OK;
  const int rc = async_wait(st);
  if (rc != SCLGPU_OK && error) *error = st->error.c_str();
  return rc;
}
void sclgpu_async_release(sclgpu_ctx* ctx) {
  AsyncState* st = nullptr;
  {
    std::lock_guard<std::mutex> lock(g_async_mu);
    auto it = g_async.find(ctx);
    if (it == g_async.end()) return;
    st = it->second;
    g_async.erase(it);
  }
  async_wait(st);
  if (st->companion) sclgpu_destroy(st->companion);
  delete st;
}

extern "C" int sclgpu_fp61_shamir_share_async(sclgpu_ctx* ctx, const uint64_t* secrets, uint64_t N, uint32_t t, uint32_t n,
                                               const uint8_t seed[16], uint64_t first_block, uint64_t* shares) {
  if (!ctx || !seed) return SCLGPU_EINVAL;
  std::vector<uint8_t> s(seed, seed + 16);  // the caller's seed buffer may go away before the worker runs
  return async_launch(ctx, [=](sclgpu_ctx* c) { return sclgpu_fp61_shamir_share(c, secrets, N, t, n, s.data(), first_block, shares); });
}
extern "C" int sclgpu_fp127_shamir_share_async(sclgpu_ctx* ctx, const void* secrets, uint64_t N, uint32_t t, uint32_t n,
                                                const uint8_t seed[16], uint64_t first_block, void* shares) {
  if (!ctx || !seed) return SCLGPU_EINVAL;
  std::vector<uint8_t> s(seed, seed + 16);
  return async_launch(ctx, [=](sclgpu_ctx* c) { return sclgpu_fp127_shamir_share(c, secrets, N, t, n, s.data(), first_block, shares); });
}
extern "C" int sclgpu_fp61_recover_p_async(sclgpu_ctx* ctx, const uint64_t* shares, uint64_t N, uint32_t n, const uint64_t* alphas,
                                            const uint64_t* x, uint64_t* out) {
  if (!ctx) return SCLGPU_EINVAL;
  std::vector<uint64_t> a;
  if (alphas) a.assign(alphas, alphas + n);
  const bool has_x = x != nullptr;
  const uint64_t xv = has_x ? *x : 0;
  return async_launch(ctx, [=](sclgpu_ctx* c) {
    return sclgpu_fp61_recover_p(c, shares, N, n, a.empty() ? nullptr : a.data(), has_x ? &xv : nullptr, out);
  });
}
extern "C" int sclgpu_fp127_recover_p_async(sclgpu_ctx* ctx, const void* shares, uint64_t N, uint32_t n, const void* alphas,
                                             const void* x, void* out) {
  if (!ctx) return SCLGPU_EINVAL;
  std::vector<uint8_t> a, xv;
  if (alphas) a.assign(static_cast<const uint8_t*>(alphas), static_cast<const uint8_t*>(alphas) + (size_t)n * 16);
  if (x) xv.assign(static_cast<const uint8_t*>(x), static_cast<const uint8_t*>(x) + 16);
  return async_launch(ctx, [=](sclgpu_ctx* c) {
    return sclgpu_fp127_recover_p(c, shares, N, n, a.empty() ? nullptr : a.data(), xv.empty() ? nullptr : xv.data(), out);
  });
}
extern "C" int sclgpu_wait(sclgpu_ctx* ctx) {
  if (!ctx) return SCLGPU_EINVAL;
  const char* err = nullptr;
  const int rc = sclgpu_async_complete(ctx, &err);
  if (rc != SCLGPU_OK && err) sclgpu_set_error(ctx, err);
  return rc;
}
