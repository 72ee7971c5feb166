// host_stage.h -- bulk transfers between PAGEABLE host memory and the device.
//
// SCL's containers are std::vector-backed (vector.h:632, matrix.h:434), so the buffers an SCL user hands
// to the host entry points are pageable.  cudaMemcpyAsync on pageable memory is staged by the driver through
// one thread (measured on the B200 box: 11 GB/s host->device, 20 GB/s device->host, against 55 GB/s for
// pinned buffers).  This stager does the same job with a ring of pinned pieces and a small team of copy
// threads, pipelined against the DMA engine:
//
//   host -> device:  [callback on `hs`: team memcpy user -> piece]  ->  [pipe stream: DMA piece -> device]
//   device -> host:  [pipe stream: DMA device -> piece]  ->  [callback on `hs`: team memcpy piece -> user]
//
// Pieces are handed round-robin; `done[s]` is recorded by the last consumer of piece s and awaited by its next
// producer, `mid[s]` orders producer -> consumer.  Host callbacks (cudaLaunchHostFunc) make no CUDA calls.
// Pinned user buffers (cudaHostAlloc / cudaHostRegister) and small transfers bypass all of this.
// Every entry point that may have staged something calls drain() before it returns.
#pragma once

#include <cuda_runtime.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace sclgpu {

// memcpy with non-temporal stores: the destination (a pinned ring piece, or the user's result buffer) is not read
// again by these threads, so write-allocate traffic and cache pollution are avoided (3 -> 2 bytes of DRAM traffic
// per byte copied).
inline void stream_copy(void* dst, const void* src, size_t bytes) {
#if defined(__x86_64__)
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
  if (bytes < 256 + head) {
    std::memcpy(d, s, bytes);
    return;
  }
  std::memcpy(d, s, head);
  d += head;
  s += head;
  bytes -= head;
  const size_t body = bytes & ~static_cast<size_t>(63);
  for (size_t i = 0; i < body; i += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
  }
  _mm_sfence();
  std::memcpy(d + body, s + body, bytes - body);
#else
  std::memcpy(dst, src, bytes);
#endif
}

class CopyTeam {
 public:
  explicit CopyTeam(int helpers) {
    try {
      for (int i = 0; i < helpers; ++i) threads_.emplace_back([this] { loop(); });
    } catch (...) {
      shutdown();
      throw;
    }
  }
  ~CopyTeam() { shutdown(); }
  // memcpy split into 1 MiB parts over the caller and the helpers; returns when every byte is copied.
  // One copy at a time (the stager's callbacks are serialised by their stream).  Job fields change only
  // while no helper is registered (active_ == 0); helpers register under the lock before reading them.
  void copy(void* dst, const void* src, size_t bytes) {
    if (bytes <= kPart || threads_.empty()) {
      stream_copy(dst, src, bytes);
      return;
    }
    {
      std::unique_lock<std::mutex> lk(m_);
      done_cv_.wait(lk, [this] { return active_ == 0; });
      dst_ = static_cast<char*>(dst);
      src_ = static_cast<const char*>(src);
      bytes_ = bytes;
      parts_ = (bytes + kPart - 1) / kPart;
      next_.store(0, std::memory_order_relaxed);
      finished_ = 0;
      ++generation_;
    }
    cv_.notify_all();
    const size_t mine = work();
    std::unique_lock<std::mutex> lk(m_);
    finished_ += mine;
    done_cv_.wait(lk, [this] { return finished_ == parts_ && active_ == 0; });
    parts_ = 0;  // a helper that wakes up late for this generation finds nothing to do
  }

 private:
  static constexpr size_t kPart = 1u << 20;
  void shutdown() {
    {
      std::lock_guard<std::mutex> lk(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
    threads_.clear();
  }
  size_t work() {
    size_t n = 0;
    for (;;) {
      const size_t p = next_.fetch_add(1, std::memory_order_relaxed);
      if (p >= parts_) return n;
      const size_t off = p * kPart, len = (bytes_ - off < kPart) ? bytes_ - off : kPart;
      stream_copy(dst_ + off, src_ + off, len);
      ++n;
    }
  }
  void loop() {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> lk(m_);
        cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
        if (stop_) return;
        seen = generation_;
        ++active_;
      }
      const size_t mine = work();
      {
        std::lock_guard<std::mutex> lk(m_);
        finished_ += mine;
        --active_;
      }
      done_cv_.notify_all();
    }
  }
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_, done_cv_;
  char* dst_ = nullptr;
  const char* src_ = nullptr;
  size_t bytes_ = 0, parts_ = 0, finished_ = 0;
  int active_ = 0;
  std::atomic<size_t> next_{0};
  uint64_t generation_ = 0;
  bool stop_ = false;
};

class HostStager {
 public:
  static constexpr size_t kPiece = 32u << 20;
  static constexpr int kSlots = 8;
  static constexpr size_t kMinStaged = 4u << 20;  // below this the driver's own path is as good

  ~HostStager() { release(); }

  // SCLGPU_HOST_STAGING=0 leaves pageable buffers to the driver (for measurement)
  static bool enabled() {
    static const bool on = [] {
      const char* v = getenv("SCLGPU_HOST_STAGING");
      return !(v && v[0] == '0');
    }();
    return on;
  }
  // true when `p` is ordinary (unregistered) host memory
  static bool pageable(const void* p) {
    if (!enabled()) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
  }

  cudaError_t h2d(cudaStream_t st, void* dev, const void* host, size_t bytes) {
    if (bytes < kMinStaged || !pageable(host) || !ready()) return cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, st);
    cudaError_t e = cudaSuccess;
    for (size_t off = 0; off < bytes && e == cudaSuccess; off += kPiece) {
      const size_t len = bytes - off < kPiece ? bytes - off : kPiece;
      const int s = next_slot();
      e = cudaStreamWaitEvent(hs_, done_[s], 0);
      if (e == cudaSuccess) e = enqueue_copy(slot_[s], static_cast<const char*>(host) + off, len);
      if (e == cudaSuccess) e = cudaEventRecord(mid_[s], hs_);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(st, mid_[s], 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(static_cast<char*>(dev) + off, slot_[s], len, cudaMemcpyHostToDevice, st);
      if (e == cudaSuccess) e = cudaEventRecord(done_[s], st);
    }
    return e;
  }

  cudaError_t d2h(cudaStream_t st, void* host, const void* dev, size_t bytes) {
    if (bytes < kMinStaged || !pageable(host) || !ready()) return cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaSuccess;
    for (size_t off = 0; off < bytes && e == cudaSuccess; off += kPiece) {
      const size_t len = bytes - off < kPiece ? bytes - off : kPiece;
      const int s = next_slot();
      e = cudaStreamWaitEvent(st, done_[s], 0);
      if (e == cudaSuccess) e = cudaMemcpyAsync(slot_[s], static_cast<const char*>(dev) + off, len, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaEventRecord(mid_[s], st);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(hs_, mid_[s], 0);
      if (e == cudaSuccess) e = enqueue_copy(static_cast<char*>(host) + off, slot_[s], len);
      if (e == cudaSuccess) e = cudaEventRecord(done_[s], hs_);
    }
    return e;
  }

  // waits for every staged copy; the task records can go afterwards
  cudaError_t drain() {
    if (!hs_ || tasks_.empty()) return cudaSuccess;
    const cudaError_t e = cudaStreamSynchronize(hs_);
    tasks_.clear();
    return e;
  }

  void release() {
    if (hs_) cudaStreamSynchronize(hs_);
    tasks_.clear();
    for (int s = 0; s < kSlots; ++s) {
      if (slot_[s]) cudaFreeHost(slot_[s]);
      if (mid_[s]) cudaEventDestroy(mid_[s]);
      if (done_[s]) cudaEventDestroy(done_[s]);
      slot_[s] = nullptr;
      mid_[s] = done_[s] = nullptr;
    }
    if (hs_) cudaStreamDestroy(hs_);
    hs_ = nullptr;
    delete team_;
    team_ = nullptr;
  }

 private:
  struct Task {
    CopyTeam* team;
    void* dst;
    const void* src;
    size_t bytes;
  };
  static void CUDART_CB run(void* p) {
    Task* t = static_cast<Task*>(p);
    t->team->copy(t->dst, t->src, t->bytes);
  }
  cudaError_t enqueue_copy(void* dst, const void* src, size_t bytes) {
    tasks_.push_back(Task{team_, dst, src, bytes});  // deque: stable addresses
    return cudaLaunchHostFunc(hs_, run, &tasks_.back());
  }
  int next_slot() {
    const int s = next_;
    next_ = (next_ + 1) % kSlots;
    return s;
  }
  // ring, events, callback stream and copy team exist (created on first use); if they cannot be created -- e.g. no
  // pinned memory left -- the transfer is left to the driver's own pageable path
  bool ready() {
    if (hs_) return true;
    if (failed_) return false;
    if (prepare() == cudaSuccess) return true;
    cudaGetLastError();
    failed_ = true;
    return false;
  }
  cudaError_t prepare() {
    if (hs_) return cudaSuccess;
    cudaError_t e = cudaStreamCreateWithFlags(&hs_, cudaStreamNonBlocking);
    for (int s = 0; s < kSlots && e == cudaSuccess; ++s) {
      e = cudaHostAlloc(&slot_[s], kPiece, cudaHostAllocDefault);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&mid_[s], cudaEventDisableTiming);
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(&done_[s], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) {
      unsigned hw = std::thread::hardware_concurrency();
      int helpers = hw >= 16 ? 7 : hw >= 8 ? 5 : hw >= 4 ? 2 : 1;  // + the callback thread; measured flat from 8 to 16 threads
      if (const char* v = getenv("SCLGPU_COPY_THREADS")) helpers = std::max(0, atoi(v) - 1);
      try {
        team_ = new CopyTeam(helpers);
      } catch (...) {  // no threads to be had: the callback thread copies alone
        team_ = new CopyTeam(0);
      }
    } else {
      release();
    }
    return e;
  }
  cudaStream_t hs_ = nullptr;
  void* slot_[kSlots] = {};
  cudaEvent_t mid_[kSlots] = {}, done_[kSlots] = {};
  int next_ = 0;
  bool failed_ = false;
  CopyTeam* team_ = nullptr;
  std::deque<Task> tasks_;
};

}  // namespace sclgpu
