// Host side of the boundary for a caller that keeps the reference's float32
// [B, N] configuration tensor (graph_builders.py:92-125) in HOST memory:
// cgsvmc_pack_configs_host converts it to the packed walker layout on the
// host cores -- 4 N bytes per walker become 8 bytes per 64 sites BEFORE the
// PCIe link, which on this pool's boxes delivers 8-25 GB/s and bounds the
// host-fed step (1.18 MB per C2 batch: 47-140 us against a 57 us kernel).
// Layout conversion only (bit = 1 <=> value > 0, exactly the rule of
// pack_kernel / load_walker); no amplitude is ever evaluated on the host.
//
// A small persistent pool (a job is a contiguous range of walkers; the caller
// takes the last share itself): workers poll briefly for the next batch of a
// feed loop before they sleep on a condition variable.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <thread>
#include <vector>

#include <immintrin.h>
#include <unistd.h>

#include "internal.h"

namespace cgsvmc {
namespace {

__attribute__((target("avx2"))) void pack_rows_avx2(const float* cfg, int64_t b0, int64_t b1, int N, int W,
                                                    uint64_t* out) {
  const __m256 zero = _mm256_setzero_ps();
  for (int64_t b = b0; b < b1; ++b) {
    const float* row = cfg + b * N;
    uint64_t* dst = out + b * W;
    for (int w = 0; w < W; ++w) {
      const int i0 = 64 * w, n = std::min(64, N - i0);
      uint64_t word = 0;
      int i = 0;
      for (; i + 8 <= n; i += 8) {
        const __m256 x = _mm256_loadu_ps(row + i0 + i);
        word |= (uint64_t)(unsigned)_mm256_movemask_ps(_mm256_cmp_ps(x, zero, _CMP_GT_OQ)) << i;
      }
      for (; i < n; ++i) word |= (uint64_t)(row[i0 + i] > 0.f) << i;
      dst[w] = word;
    }
  }
}

void pack_rows_scalar(const float* cfg, int64_t b0, int64_t b1, int N, int W, uint64_t* out) {
  for (int64_t b = b0; b < b1; ++b) {
    const float* row = cfg + b * N;
    uint64_t* dst = out + b * W;
    for (int w = 0; w < W; ++w) {
      const int i0 = 64 * w, n = std::min(64, N - i0);
      uint64_t word = 0;
      for (int i = 0; i < n; ++i) word |= (uint64_t)(row[i0 + i] > 0.f) << i;
      dst[w] = word;
    }
  }
}

void pack_rows(const float* cfg, int64_t b0, int64_t b1, int N, int W, uint64_t* out) {
  static const bool avx2 = __builtin_cpu_supports("avx2");
  if (avx2) pack_rows_avx2(cfg, b0, b1, N, W, out);
  else pack_rows_scalar(cfg, b0, b1, N, W, out);
}

struct Job {
  const float* cfg = nullptr;
  uint64_t* out = nullptr;
  int64_t B = 0;
  int N = 0, W = 0, shares = 1;
};

class Pool {
 public:
  // one caller at a time (the pool is process-wide)
  void run(const Job& job) {
    std::lock_guard<std::mutex> call(call_mu_);
    const int helpers = job.shares - 1;
    if (owner_pid_ != getpid()) {      // after fork() the child has the bookkeeping but not the threads
      workers_.clear();
      owner_pid_ = getpid();
    }
    if (helpers > 0) {
      while ((int)workers_.size() < helpers) {
        const int id = (int)workers_.size();
        workers_.emplace_back([this, id] { loop(id); });
        workers_.back().detach();        // never joined: the pool lives as long as the process
      }
      job_ = job;
      // EVERY worker acknowledges every job (after it has read job_), also those without a
      // share: the next call must not overwrite job_ under a late reader
      pending_.store((int)workers_.size(), std::memory_order_relaxed);
      {
        std::lock_guard<std::mutex> lk(mu_);      // (a worker about to sleep re-checks under this lock)
        generation_.fetch_add(1, std::memory_order_release);
      }
      cv_.notify_all();
    }
    share(job, job.shares - 1);
    while (pending_.load(std::memory_order_acquire) != 0) _mm_pause();
  }

 private:
  static void share(const Job& j, int k) {
    const int64_t per = (j.B + j.shares - 1) / j.shares;
    const int64_t b0 = std::min<int64_t>(j.B, per * k), b1 = std::min<int64_t>(j.B, b0 + per);
    if (b1 > b0) pack_rows(j.cfg, b0, b1, j.N, j.W, j.out);
  }

  // A feed loop submits a batch every ~70 us: a worker that has just finished
  // keeps polling for the next job for kSpinUs before it goes to sleep on the
  // condition variable (a futex wake-up costs 5-10 us per worker -- more than
  // its share of a C2 batch).
  void loop(int id) {
    constexpr int64_t kSpinUs = 300;
    uint64_t seen = 0;
    for (;;) {
      const auto t0 = std::chrono::steady_clock::now();
      bool got = false;
      for (int it = 0;; ++it) {
        if (generation_.load(std::memory_order_acquire) != seen) { got = true; break; }
        _mm_pause();
        if ((it & 63) == 63 &&
            std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count() > kSpinUs)
          break;
      }
      if (!got) {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return generation_.load(std::memory_order_acquire) != seen; });
      }
      seen = generation_.load(std::memory_order_acquire);
      const Job j = job_;
      if (id < j.shares - 1) share(j, id);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }

  std::mutex mu_, call_mu_;
  std::condition_variable cv_;
  std::vector<std::thread> workers_;
  Job job_;
  std::atomic<uint64_t> generation_{0};
  std::atomic<int> pending_{0};
  pid_t owner_pid_ = getpid();
};

Pool& pool() {
  static Pool* p = new Pool();      // never destroyed: worker threads must not be joined at exit of a host process
  return *p;
}

}  // namespace

int pack_configs_host(const float* configs, int64_t B, int N, uint64_t* packed, int n_threads) {
  Job j;
  j.cfg = configs; j.out = packed; j.B = B; j.N = N; j.W = n_words(N);
  const int64_t bytes = B * (int64_t)N * 4;
  int shares = n_threads > 0 ? n_threads : (int)std::min<int64_t>(8, std::max<int64_t>(1, bytes / (128 << 10)));
  const unsigned hw = std::thread::hardware_concurrency();
  if (hw > 0) shares = std::min<int>(shares, (int)hw);
  j.shares = std::max(1, std::min<int>(shares, 64));
  pool().run(j);
  return CGSVMC_OK;
}

}  // namespace cgsvmc
