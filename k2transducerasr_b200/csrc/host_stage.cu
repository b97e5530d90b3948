// Host-to-device copies from PAGEABLE memory, staged by the library itself.
//
// The reference's seam hands over managed arrays (float[] encoder_out, ref Model/EncoderOutputEntity.cs:10-20): pinned by the
// runtime for the duration of a P/Invoke call, but not page-locked. cudaMemcpyAsync from such memory goes through the driver's
// single bounce buffer with one memcpy thread: ~11-13 GB/s on the B200 boxes against ~55 GB/s for page-locked memory (measured:
// bench.py e2e_pageable vs e2e). Here the same staging is done with several host threads copying into two page-locked bounce
// buffers while the DMA of the previous slice is in flight. Callers that can should still take their buffers from k2b_host_alloc
// (or register them): that path needs no host copy at all.
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "k2b_internal.h"

namespace k2b {

struct HostStage {
  static constexpr size_t kBounce = size_t(16) << 20;
  char* bounce[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  bool busy[2] = {false, false};
  // a tiny persistent pool: the handle is single-threaded towards its caller, these threads only ever run memcpy
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::function<void(int, int)> job;      // (worker index, worker count)
  int generation = 0, pending = 0;
  bool stop = false;

  void worker(int idx, int n) {
    int seen = 0;
    for (;;) {
      std::function<void(int, int)> f;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
        f = job;
      }
      f(idx, n);
      {
        std::lock_guard<std::mutex> lk(mu);
        if (--pending == 0) cv_done.notify_all();
      }
    }
  }
  void start(int n) {
    for (int i = 0; i < n; ++i) threads.emplace_back([this, i, n] { worker(i, n); });
  }
  void run(const std::function<void(int, int)>& f) {
    if (threads.empty()) { f(0, 1); return; }
    std::unique_lock<std::mutex> lk(mu);
    job = f;
    pending = (int)threads.size();
    ++generation;
    cv_work.notify_all();
    cv_done.wait(lk, [&] { return pending == 0; });
  }
  ~HostStage() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv_work.notify_all();
    for (auto& t : threads) t.join();
    for (int i = 0; i < 2; ++i) {
      if (done[i]) cudaEventDestroy(done[i]);
      if (bounce[i]) cudaFreeHost(bounce[i]);
    }
  }
};

void host_stage_free(k2b_handle* h) {
  delete h->host_stage;
  h->host_stage = nullptr;
}

static bool is_pageable(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
  return a.type == cudaMemoryTypeUnregistered;
}

static int32_t ensure_stage(k2b_handle* h) {
  if (h->host_stage != nullptr) return K2B_OK;
  HostStage* s = new HostStage();
  h->host_stage = s;
  for (int i = 0; i < 2; ++i) {
    K2B_CUDA(h, cudaHostAlloc(reinterpret_cast<void**>(&s->bounce[i]), HostStage::kBounce, cudaHostAllocDefault));
    K2B_CUDA(h, cudaEventCreateWithFlags(&s->done[i], cudaEventDisableTiming));
  }
  unsigned hw = std::thread::hardware_concurrency();
  int n = (int)(hw / 4);                                        // copy threads: a quarter of the host's (several ranks share it)
  n = n > 8 ? 8 : (n < 2 ? (hw >= 4 ? 2 : 0) : n);
  if (h->opt_copy_threads >= 0) n = h->opt_copy_threads;
  s->start(n);
  return K2B_OK;
}

// dst (device, pitch dpitch) <- src (host, pitch spitch): `rows` rows of `width` bytes, enqueued on `stream`. Page-locked sources
// go straight to cudaMemcpy2DAsync; pageable ones are staged slice by slice through the two bounce buffers. Returns once the
// last slice is enqueued (the caller's buffer is no longer read after that).
int32_t h2d_rows(k2b_handle* h, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows, cudaStream_t stream) {
  if (rows == 0 || width == 0) return K2B_OK;
  if (!is_pageable(src)) {
    if (rows == 1) K2B_CUDA(h, cudaMemcpyAsync(dst, src, width, cudaMemcpyHostToDevice, stream));
    else K2B_CUDA(h, cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, cudaMemcpyHostToDevice, stream));
    return K2B_OK;
  }
  K2B_TRY(ensure_stage(h));
  HostStage* s = h->host_stage;
  int k = 0;
  auto acquire = [&](int i) -> int32_t {
    if (s->busy[i]) { K2B_CUDA(h, cudaEventSynchronize(s->done[i])); s->busy[i] = false; }
    return K2B_OK;
  };
  const char* sp = static_cast<const char*>(src);
  char* dp = static_cast<char*>(dst);
  if (width <= HostStage::kBounce / 2 && rows > 1) {
    const size_t per = HostStage::kBounce / width;                    // rows per slice
    for (size_t r0 = 0; r0 < rows; r0 += per, k ^= 1) {
      const size_t nr = rows - r0 < per ? rows - r0 : per;
      K2B_TRY(acquire(k));
      char* b = s->bounce[k];
      s->run([&](int idx, int n) {
        for (size_t r = (size_t)idx; r < nr; r += (size_t)n) memcpy(b + r * width, sp + (r0 + r) * spitch, width);
      });
      K2B_CUDA(h, cudaMemcpy2DAsync(dp + r0 * dpitch, dpitch, b, width, width, nr, cudaMemcpyHostToDevice, stream));
      K2B_CUDA(h, cudaEventRecord(s->done[k], stream));
      s->busy[k] = true;
    }
    return K2B_OK;
  }
  for (size_t r = 0; r < rows; ++r) {                                 // long rows (or one flat range): slices of bytes
    for (size_t o = 0; o < width; o += HostStage::kBounce, k ^= 1) {
      const size_t nb = width - o < HostStage::kBounce ? width - o : HostStage::kBounce;
      K2B_TRY(acquire(k));
      char* b = s->bounce[k];
      const char* from = sp + r * spitch + o;
      s->run([&](int idx, int n) {
        const size_t piece = ((nb + n - 1) / n + 63) & ~size_t(63);
        const size_t lo = piece * (size_t)idx;
        if (lo < nb) memcpy(b + lo, from + lo, nb - lo < piece ? nb - lo : piece);
      });
      K2B_CUDA(h, cudaMemcpyAsync(dp + r * dpitch + o, b, nb, cudaMemcpyHostToDevice, stream));
      K2B_CUDA(h, cudaEventRecord(s->done[k], stream));
      s->busy[k] = true;
    }
  }
  return K2B_OK;
}

}  // namespace k2b
