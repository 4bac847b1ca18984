// hostpipe.hpp -- host-side plumbing of the engine: helper-thread teams for packing ForceBatch buffers
// into pinned staging, CPU placement next to a GPU, and one persistent worker thread per device for
// batches that are cut across several GPUs ("one host thread + streams per GPU, pinned
// staging, D2H directly into the caller's out[i].F, then a host join").
//
// Nothing here computes: the reference's batch contract (CppCore/rgpot/Potential.hpp:225-240,
// ForceStructs.hpp:40-54) hands over nSystems independent, separately allocated, pageable buffers;
// a GPU can only DMA from page-locked memory, so the bytes have to be moved once on the host, and at
// 1e9 atom-evals/s one core's memcpy (343 MB each way per 65 536-configuration batch) would cost six
// times what the kernel does.
#pragma once

#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <string.h>
#include <unistd.h>

#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace rb {

// CPUs this process may run on
inline int usable_cpus() {
  cpu_set_t set;
  CPU_ZERO(&set);
  if (sched_getaffinity(0, sizeof set, &set) == 0) {
    const int c = CPU_COUNT(&set);
    if (c > 0) return c;
  }
  const unsigned hc = std::thread::hardware_concurrency();
  return hc ? (int)hc : 1;
}

// The CPUs next to a GPU (its NUMA node), restricted to what `allowed` permits.  Empty set when the
// topology cannot be read (containers without sysfs): the caller then leaves placement alone.
inline bool gpu_local_cpus(int device, const cpu_set_t& allowed, cpu_set_t& out) {
  CPU_ZERO(&out);
  char bdf[32] = {0};
  if (cudaDeviceGetPCIBusId(bdf, sizeof bdf, device) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  for (char* p = bdf; *p; ++p)
    if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
  char path[128];
  snprintf(path, sizeof path, "/sys/bus/pci/devices/%s/local_cpulist", bdf);
  FILE* f = fopen(path, "r");
  if (!f) return false;
  char line[4096] = {0};
  const bool got = fgets(line, sizeof line, f) != nullptr;
  fclose(f);
  if (!got) return false;
  int any = 0;
  for (char* tok = strtok(line, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
    int lo = 0, hi = 0;
    const int k = sscanf(tok, "%d-%d", &lo, &hi);
    if (k < 1) continue;
    if (k == 1) hi = lo;
    for (int c = lo; c <= hi && c < CPU_SETSIZE; ++c)
      if (CPU_ISSET(c, &allowed)) {
        CPU_SET(c, &out);
        ++any;
      }
  }
  return any > 0;
}

// A fixed team of helper threads; run(n, fn) calls fn(begin, end) over [0, n) in dynamic slices on
// the helpers AND the calling thread, and returns when all of it is done.  Between jobs a helper
// spins for a short while before it goes to sleep: the batch pipeline issues a job every few hundred
// microseconds, and a futex wake-up per job and helper would cost more than the copies themselves.
class Team {
 public:
  explicit Team(int helpers, const cpu_set_t* cpus = nullptr) {
    if (cpus) {
      cpus_ = *cpus;
      pinned_ = true;
    }
    for (int t = 0; t < helpers; ++t) th_.emplace_back([this] { loop(); });
  }
  ~Team() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_.store(true, std::memory_order_release);
    }
    cv_.notify_all();
    for (auto& t : th_) t.join();
  }
  Team(const Team&) = delete;
  Team& operator=(const Team&) = delete;
  int size() const { return (int)th_.size() + 1; }

  template <class F>
  void run(size_t n, size_t grain, F&& fn) {
    if (n == 0) return;
    if (th_.empty() || n <= grain) {
      fn((size_t)0, n);
      return;
    }
    std::function<void(size_t, size_t)> f = std::forward<F>(fn);
    job_ = &f;
    n_ = n;
    grain_ = grain ? grain : 1;
    next_.store(0, std::memory_order_relaxed);
    pending_.store((int)th_.size(), std::memory_order_relaxed);
    {
      std::lock_guard<std::mutex> lk(mu_);  // (pairs with the sleepers' predicate check)
      gen_.fetch_add(1, std::memory_order_release);
    }
    if (sleepers_.load(std::memory_order_acquire) > 0) cv_.notify_all();
    work();
    // the helpers finish within microseconds of the caller: spin, do not sleep
    while (pending_.load(std::memory_order_acquire) != 0) cpu_relax();
    job_ = nullptr;
  }

 private:
  static void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
  }
  void work() {
    for (;;) {
      const size_t b = next_.fetch_add(grain_, std::memory_order_relaxed);
      if (b >= n_) break;
      (*job_)(b, b + grain_ < n_ ? b + grain_ : n_);
    }
  }
  void loop() {
    if (pinned_) sched_setaffinity(0, sizeof cpus_, &cpus_);
    unsigned long seen = 0;
    for (;;) {
      // spin for about 200 microseconds, then sleep until the next job is announced
      bool got = false;
      for (int spin = 0; spin < 20000; ++spin) {
        if (gen_.load(std::memory_order_acquire) != seen || stop_.load(std::memory_order_acquire)) {
          got = true;
          break;
        }
        cpu_relax();
      }
      if (!got) {
        std::unique_lock<std::mutex> lk(mu_);
        sleepers_.fetch_add(1, std::memory_order_release);
        cv_.wait(lk, [&] { return stop_.load(std::memory_order_acquire) || gen_.load(std::memory_order_acquire) != seen; });
        sleepers_.fetch_sub(1, std::memory_order_release);
      }
      if (stop_.load(std::memory_order_acquire)) return;
      seen = gen_.load(std::memory_order_acquire);
      work();
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  std::vector<std::thread> th_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::function<void(size_t, size_t)>* job_ = nullptr;
  std::atomic<size_t> next_{0};
  size_t n_ = 0, grain_ = 1;
  std::atomic<unsigned long> gen_{0};
  std::atomic<int> pending_{0};
  std::atomic<int> sleepers_{0};
  std::atomic<bool> stop_{false};
  bool pinned_ = false;
  cpu_set_t cpus_;
};

// One persistent thread that executes closures in order (a device's worker: its CUDA calls, its
// pinned allocations and its helper team stay on the CPUs next to that device).
class Worker {
 public:
  explicit Worker(const cpu_set_t* cpus = nullptr) {
    if (cpus) {
      cpus_ = *cpus;
      pinned_ = true;
    }
    th_ = std::thread([this] { loop(); });
  }
  ~Worker() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    th_.join();
  }
  Worker(const Worker&) = delete;
  Worker& operator=(const Worker&) = delete;
  void submit(std::function<void()> f) {
    {
      std::lock_guard<std::mutex> lk(mu_);
      job_ = std::move(f);
      has_ = true;
      busy_ = true;
    }
    cv_.notify_all();
  }
  void wait() {
    std::unique_lock<std::mutex> lk(mu_);
    done_.wait(lk, [this] { return !busy_; });
  }

 private:
  void loop() {
    if (pinned_) sched_setaffinity(0, sizeof cpus_, &cpus_);
    for (;;) {
      std::function<void()> f;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || has_; });
        if (stop_) return;
        f = std::move(job_);
        has_ = false;
      }
      f();
      {
        std::lock_guard<std::mutex> lk(mu_);
        busy_ = false;
      }
      done_.notify_all();
    }
  }
  std::thread th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::function<void()> job_;
  bool has_ = false, busy_ = false, stop_ = false, pinned_ = false;
  cpu_set_t cpus_;
};

// memcpy whose stores bypass the cache (SSE2 streaming stores, part of baseline x86-64).  Packing a
// batch moves hundreds of megabytes that the CPU never reads again: an ordinary store first reads the
// destination line into the cache (read for ownership), so the copy costs three memory transfers per
// byte where a streaming store costs two -- and the host's memory system, shared with the GPUs' DMA,
// is what bounds the pageable entry points.  Call copy_stream_fence() before anything else (a DMA
// engine, another thread) reads the destination.
inline void copy_stream(void* dst, const void* src, size_t n) {
#if defined(__x86_64__)
  char* d = static_cast<char*>(dst);
  const char* s = static_cast<const char*>(src);
  if (n < 256) {
    memcpy(d, s, n);
    return;
  }
  size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
  if (head) {
    memcpy(d, s, head);
    d += head;
    s += head;
    n -= head;
  }
  size_t blocks = n / 64;
  for (; blocks; --blocks, d += 64, s += 64) {
    const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s));
    const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 16));
    const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 32));
    const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + 48));
    _mm_stream_si128(reinterpret_cast<__m128i*>(d), a);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 16), b);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 32), c);
    _mm_stream_si128(reinterpret_cast<__m128i*>(d + 48), e);
  }
  n &= 63;
  if (n) memcpy(d, s, n);
#else
  memcpy(dst, src, n);
#endif
}
inline void copy_stream_fence() {
#if defined(__x86_64__)
  _mm_sfence();
#endif
}

// Is this host pointer page-locked (cudaMallocHost / cudaHostRegister)?  Pageable memory cannot be
// DMA'd: it goes through pinned staging.
inline bool host_ptr_is_pinned(const void* p) {
  if (!p) return false;
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeHost;
}

}  // namespace rb
