// Error reporting, per-stage CUDA-event profiling, the device buffer pool.
// Part of the single translation unit plspm_b200.cu (included there, in this order); see DESIGN.md §4.
#pragma once

// ------------------------------------------------------------------------------------------------
// error handling / profiling
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? PLSPM_ERR_NOMEM : PLSPM_ERR_CUDA,              \
                  std::string(#expr) + ": " + cudaGetErrorString(e_));                             \
  } while (0)

enum { ST_COUNTS = 0, ST_GRAM = 1, ST_REDUCE = 2, ST_SOLVE = 3, ST_SCORES = 4, ST_UPLOAD = 5, ST_COLSUM = 6, ST_CROSS = 7, ST_SCOREGEN = 8, ST_CONV = 9, ST_GRAM_I8 = 10, ST_FINALIZE = 11, ST_N = 12 };
struct Profile {
  std::mutex mu;
  double ms[ST_N] = {0};
  int64_t launches[ST_N] = {0};
};
static Profile g_prof;
static std::atomic<int64_t> g_redo_count{0};  // replicates redone exactly after an undecided low-precision vote

// Process-wide cache of large device buffers (per device): plspm_bootstrap_host() creates and
// destroys a data handle per call, and cudaMalloc/cudaFree of the 0.2-1.5 GB buffers would
// otherwise dominate the end-to-end time of a call.
struct DevPool {
  struct Block { void* p; size_t bytes; int device; };
  std::mutex mu;
  std::vector<Block> free_blocks;
  size_t cached = 0;
  // Released buffers are kept for the next call up to this many bytes (PLSPM_POOL_GB, default 8: enough for the
  // data handle + workspace of a c3-sized plspm_bootstrap_host call); plspm_pool_set_limit / plspm_pool_trim
  // let a host that shares the device with other CUDA users (torch, NCCL) lower it or give everything back.
  size_t max_cached = (size_t)((getenv("PLSPM_POOL_GB") ? atof(getenv("PLSPM_POOL_GB")) : 8.0) * (double)((size_t)1 << 30));
  void set_limit(size_t bytes) {
    {
      std::lock_guard<std::mutex> lk(mu);
      max_cached = bytes;
    }
    if (cached > bytes) trim();
  }
  // Requests are rounded up to a size class and only an exact class match is reused: the allocation pattern
  // of a call (data handle + workspace) is deterministic, so from the second identical call on every request
  // hits the cache -- a best-fit policy kept trading blocks between requests for several calls.
  static size_t size_class(size_t bytes) {
    const size_t g = bytes <= ((size_t)64 << 10) ? 256 : bytes <= ((size_t)16 << 20) ? ((size_t)64 << 10) : ((size_t)2 << 20);
    return (bytes + g - 1) / g * g;
  }
  cudaError_t alloc(void** out, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    bytes = size_class(std::max<size_t>(bytes, 1));
    {
      std::lock_guard<std::mutex> lk(mu);
      for (int i = (int)free_blocks.size() - 1; i >= 0; --i) {
        const Block& b = free_blocks[i];
        if (b.device == dev && b.bytes == bytes) {
          *out = b.p;
          cached -= b.bytes;
          sizes.push_back({*out, b.bytes, dev});
          free_blocks.erase(free_blocks.begin() + i);
          return cudaSuccess;
        }
      }
    }
    ++misses;
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) {  // release the cache and retry once
      trim();
      e = cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) {
      std::lock_guard<std::mutex> lk(mu);
      sizes.push_back({*out, bytes, dev});
    }
    return e;
  }
  void release(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> lk(mu);
    for (size_t i = 0; i < sizes.size(); ++i)
      if (sizes[i].p == p) {
        Block b = sizes[i];
        sizes.erase(sizes.begin() + i);
        if (cached + b.bytes <= max_cached) {
          free_blocks.push_back(b);
          cached += b.bytes;
        } else {
          ++evictions;
          cudaFree(p);
        }
        return;
      }
    cudaFree(p);
  }
  void trim() {
    std::lock_guard<std::mutex> lk(mu);
    for (auto& b : free_blocks) cudaFree(b.p);
    free_blocks.clear();
    cached = 0;
  }
  std::vector<Block> sizes;  // live blocks handed out
  std::atomic<int64_t> misses{0}, evictions{0};  // cudaMalloc / cudaFree calls the cache did not absorb (PLSPM_TRACE)
};
static DevPool g_pool;

// A timed launch region: events on the launching stream; durations are collected when the
// stream is synchronised at the end of the API call.
struct StageTimer {
  struct Rec { int stage; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
  void begin(int stage, cudaStream_t s) {
    Rec r{stage, get(), get()};
    cudaEventRecord(r.a, s);
    recs.push_back(r);
  }
  void end(cudaStream_t s) { cudaEventRecord(recs.back().b, s); }
  // call after the stream is synchronised; `upto` < recs.size(): only the first `upto` records are known to have
  // completed (a later batch is already in flight), the rest stay queued
  void collect(size_t upto = (size_t)-1) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    const size_t n = std::min(upto, recs.size());
    for (size_t i = 0; i < n; ++i) {
      Rec& r = recs[i];
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) g_prof.ms[r.stage] += ms;
      g_prof.launches[r.stage] += 1;
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    recs.erase(recs.begin(), recs.begin() + n);
  }
  ~StageTimer() {
    for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : pool) cudaEventDestroy(e);
  }
};
