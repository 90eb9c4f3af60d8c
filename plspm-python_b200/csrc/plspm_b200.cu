// plspm_b200: sm_100a kernels + C ABI (include/plspm_b200.h).
//
// Data flow of one bootstrap batch (replicates are the batch dimension):
//
//   counts_kernel   Philox4x32-10 (or injected) resample indices -> multiplicity c[b][i] (u32)
//   gram_kernel     weighted second moments per replicate, fp64:
//                     G_b[p,q] = sum_i c_bi x~_ip x~_iq  (8x8 register tiles),  colsum_b[p]
//                   X~ row tiles are staged global->shared by the TMA engine (cp.async.bulk +
//                   mbarrier ring, one producer warp) and shared by every (replicate, tile
//                   group) warp of the CTA: X is read from HBM once per wave of replicates,
//                   not once per replicate and iteration.
//   reduce_kernel   fixed-order sum over row chunks (only when rows are split, e.g. one fit)
//   solve_kernel    one CTA per replicate: the whole PLS-PM iteration in the covariance
//                   domain (solver_core.h), inner model, effects, loadings
//   scores_kernel   single fit only: scores = X~ . coef - shift   (N x L, HBM-bound)
//
// Nothing here falls back to a CPU path: without a CUDA device every entry point fails.
#include <cublas_v2.h>
#include <type_traits>
#include <chrono>
#include <map>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/plspm_b200.h"
#include "plspm_model.h"
#include "solver_core.h"
#include "solver_num.h"

using namespace plspm;

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

enum { ST_COUNTS = 0, ST_GRAM = 1, ST_REDUCE = 2, ST_SOLVE = 3, ST_SCORES = 4, ST_UPLOAD = 5, ST_COLSUM = 6, ST_CROSS = 7, ST_SCOREGEN = 8, ST_CONV = 9, ST_GRAM_I8 = 10, ST_N = 12 };
struct Profile {
  std::mutex mu;
  double ms[ST_N] = {0};
  int64_t launches[ST_N] = {0};
};
static Profile g_prof;
static int64_t g_redo_count = 0;  // replicates redone exactly after an undecided low-precision vote

// Process-wide cache of large device buffers (per device): plspm_bootstrap_host() creates and
// destroys a data handle per call, and cudaMalloc/cudaFree of the 0.2-1.5 GB buffers would
// otherwise dominate the end-to-end time of a call.
struct DevPool {
  struct Block { void* p; size_t bytes; int device; };
  std::mutex mu;
  std::vector<Block> free_blocks;
  size_t cached = 0;
  static constexpr size_t kMaxCached = (size_t)64 << 30;
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
        if (cached + b.bytes <= kMaxCached) {
          free_blocks.push_back(b);
          cached += b.bytes;
        } else {
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
  void collect() {  // call after the stream is synchronised
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (auto& r : recs) {
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) g_prof.ms[r.stage] += ms;
      g_prof.launches[r.stage] += 1;
      pool.push_back(r.a);
      pool.push_back(r.b);
    }
    recs.clear();
  }
  ~StageTimer() {
    for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : pool) cudaEventDestroy(e);
  }
};

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct plspm_model {
  HostModel h;
  int device = 0;
  bool numeric = false;  // non-metric treatment with numeric scales (plspm_model_set_numeric)
  std::vector<void*> dev_allocs;
  ModelView dv;  // device pointers
};

// cuBLAS handles, recycled per device (a data handle owns one exclusively while it lives)
static std::mutex g_blas_mu;
static std::map<int, std::vector<cublasHandle_t>> g_blas_free;
static cublasHandle_t blas_acquire(int device) {
  {
    std::lock_guard<std::mutex> lk(g_blas_mu);
    auto& v = g_blas_free[device];
    if (!v.empty()) {
      cublasHandle_t h = v.back();
      v.pop_back();
      return h;
    }
  }
  cublasHandle_t h = nullptr;
  return cublasCreate(&h) == CUBLAS_STATUS_SUCCESS ? h : nullptr;
}
static void blas_release(int device, cublasHandle_t h) {
  std::lock_guard<std::mutex> lk(g_blas_mu);
  g_blas_free[device].push_back(h);
}

struct Workspace {
  void* ptr = nullptr;
  size_t bytes = 0;
};

struct plspm_data {
  const plspm_model* model = nullptr;
  int64_t N = 0;
  double* X = nullptr;   // [N][Ppad] slot layout, globally centred
  double* mu = nullptr;  // [Ppad]
  // low-precision copy for the tensor-core sign vote (sparse tile sets): xh = x~ / sd (fp16)
  __half* Xh = nullptr;      // [N][Ppad]
  float* Xf = nullptr;       // [N][Ppad] fp32 copy of x~ (score generation of the sign vote)
  double* inv_sd = nullptr;  // [Ppad]
  cublasHandle_t blas = nullptr;
  int blas_device = 0;
  bool fast_vote = false;    // the fp16 pass is worth trying on this data
  // exact integer digit planes of x~ for the tensor-core column sums (see digits_kernel)
  int8_t* D8 = nullptr;      // [I8_DIGITS * Ppad][Npad]
  double* dscale = nullptr;  // [Ppad] value of one unit of the least significant digit
  int64_t Npad = 0;
  bool i8_colsum = false;
  // ... and of the pair products z_i = x~_ip x~_iq of the model's Gram tile set (see zdigits_kernel)
  int8_t* Z8 = nullptr;      // [I8_DIGITS * n_zcols][Npad] resident planes, or null: generated per row chunk per batch
  int *zp = nullptr, *zq = nullptr;       // [n_zcols] the two columns of every pair
  double* zqscale = nullptr;              // [n_zcols] 2^(40 - e_p - e_q)
  int64_t z_chunk_rows = 0;               // streaming mode: rows per generated chunk
  double* zdscale = nullptr; // [n_zcols]
  int *zdst = nullptr, *zdst2 = nullptr;  // [n_zcols] offsets into a replicate's tile array (mirror or -1)
  int n_zcols = 0;
  const plspm_model* z_model = nullptr;   // the tile set the planes were built for
  cudaStream_t stream = nullptr;
  Workspace ws;          // grown on demand, reused across calls
  StageTimer timer;
  int sm_count = 148;
  int max_smem = 227 * 1024;
};

template <class T>
static int upload_vec(plspm_model* m, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CK(cudaMalloc(&p, bytes));
  m->dev_allocs.push_back(p);
  if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)p;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA engine, 1-D)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Random123); counter = (row group, 0, replicate lo, replicate hi), key = seed
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                      uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__host__ __device__ __forceinline__ uint32_t index_from_u32(uint32_t u, uint32_t N) {
  return (uint32_t)(((uint64_t)u * (uint64_t)N) >> 32);
}

// counts[b][i] += multiplicity of row i in replicate b.  One thread = 4 consecutive draws.
__global__ void counts_kernel(uint32_t* __restrict__ counts, const int32_t* __restrict__ idx, int64_t N, int64_t nrep,
                              int64_t rep_begin, uint64_t seed) {
  const int64_t groups = (N + 3) / 4;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= groups * nrep) return;
  const int64_t b = gid / groups, g = gid - b * groups;
  uint32_t* c = counts + b * N;
  if (idx) {
    const int32_t* ib = idx + b * N;
    for (int k = 0; k < 4; ++k) {
      int64_t i = g * 4 + k;
      if (i < N) atomicAdd(&c[(uint32_t)ib[i]], 1u);
    }
  } else {
    uint64_t rep = (uint64_t)(rep_begin + b);
    uint32_t r[4];
    philox4x32_10((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    for (int k = 0; k < 4; ++k)
      if (g * 4 + k < N) atomicAdd(&c[index_from_u32(r[k], (uint32_t)N)], 1u);
  }
}

__global__ void indices_kernel(int32_t* __restrict__ out, int64_t N, uint64_t rep, uint64_t seed) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (N + 3) / 4) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  for (int k = 0; k < 4; ++k)
    if (g * 4 + k < N) out[g * 4 + k] = (int32_t)index_from_u32(r[k], (uint32_t)N);
}

// ------------------------------------------------------------------------------------------------
// upload: column means (two-stage, fixed order) and slot-layout relayout with centring
// ------------------------------------------------------------------------------------------------
__global__ void colsum_partial_kernel(const double* __restrict__ X, int64_t N, int64_t ld, int P, int64_t rows_per_block,
                                      double* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, N);
  for (int p = threadIdx.x; p < P; p += blockDim.x) {
    double s = 0.0;
    for (int64_t i = r0; i < r1; ++i) s += X[i * ld + p];
    partial[(int64_t)blockIdx.x * P + p] = s;
  }
}
__global__ void colmean_final_kernel(const double* __restrict__ partial, int nblocks, int P, int64_t N,
                                     const int* __restrict__ src_col, double* __restrict__ mu) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * P + p];
  mu[src_col[p]] = s / (double)N;
}
__global__ void relayout_kernel(const double* __restrict__ X, int64_t N, int64_t ld, int Ppad,
                                const int* __restrict__ col_src, const double* __restrict__ mu,
                                double* __restrict__ out) {
  const int64_t total = N * Ppad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = e / Ppad;
    int c = (int)(e - i * Ppad);
    int s = col_src[c];
    out[e] = (s >= 0) ? X[i * ld + s] - mu[c] : 0.0;
  }
}

// column sums of squares of the centred slot-layout matrix -> 1/sd, and the fp16 copy xh = x~/sd
__global__ void colsq_partial_kernel(const double* __restrict__ X, int64_t N, int Ppad, int64_t rows_per_block,
                                     double* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, N);
  for (int p = threadIdx.x; p < Ppad; p += blockDim.x) {
    double s = 0.0;
    for (int64_t i = r0; i < r1; ++i) s = fma(X[i * Ppad + p], X[i * Ppad + p], s);
    partial[(int64_t)blockIdx.x * Ppad + p] = s;
  }
}
__global__ void inv_sd_kernel(const double* __restrict__ partial, int nblocks, int Ppad, int64_t N,
                              double* __restrict__ inv_sd) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ppad) return;
  double s = 0.0;
  for (int b = 0; b < nblocks; ++b) s += partial[(int64_t)b * Ppad + p];
  inv_sd[p] = s > 0.0 ? 1.0 / sqrt(s / (double)N) : 0.0;
}
__global__ void make_half_kernel(const double* __restrict__ X, int64_t N, int Ppad, const double* __restrict__ inv_sd,
                                 __half* __restrict__ out) {
  const int64_t total = N * Ppad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = __double2half(X[e] * inv_sd[e % Ppad]);
}

// ---- column sums on the tensor cores, exactly ------------------------------------------------------
// colsum[b][p] = sum_i c_bi x~_ip has a small-integer operand (the multiplicities), so it can be an INT8
// GEMM with int32 accumulation -- exact integer arithmetic -- if x~ is an integer too.  At upload every
// column is scaled by a power of two to |q| <= 2^40 (q = rint(x~ 2^(40-e_p)), 2^e_p >= max|x~_p|) and q is
// split into six balanced base-128 digits d_k in [-64, 63], stored as int8 planes D8[k][p][i] (k-major,
// each column contiguous over the rows: the "TN" operand layout of the IMMA kernels).  Per batch:
// S_k = counts8 x D8_k (one cuBLAS int8 GEMM over all planes), colsum = dscale_p * sum_k 128^k S_k.
// Rounding: |x~ - q 2^(e_p-40)| <= 2^(e_p-41), i.e. 4.5e-13 of the column's largest value, random in sign.
// Multiplicities above 127 (impossible for practical bootstrap draws, possible with injected indices)
// raise a flag and the batch is redone with the fp64 kernel.
constexpr int I8_DIGITS = 6;
constexpr int64_t I8_KCHUNK = 262144;  // rows per GEMM: 262144 * 64 * 127 < 2^31
__global__ void colabsmax_partial_kernel(const double* __restrict__ X, int64_t N, int Ppad, int64_t rows_per_block,
                                         double* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, N);
  for (int p = threadIdx.x; p < Ppad; p += blockDim.x) {
    double m = 0.0;
    for (int64_t i = r0; i < r1; ++i) m = fmax(m, fabs(X[i * Ppad + p]));
    partial[(int64_t)blockIdx.x * Ppad + p] = m;
  }
}
__global__ void digit_scale_kernel(const double* __restrict__ partial, int nblocks, int Ppad, double* __restrict__ dscale,
                                   double* __restrict__ qscale) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ppad) return;
  double m = 0.0;
  for (int b = 0; b < nblocks; ++b) m = fmax(m, partial[(int64_t)b * Ppad + p]);
  int e = 0;
  if (m > 0.0) {
    frexp(m, &e);  // m = f 2^e, f in [0.5, 1): 2^e > m
  }
  dscale[p] = ldexp(1.0, e - 40);
  qscale[p] = ldexp(1.0, 40 - e);
}
// Balanced base-128 digits without carries: with U = q + sum_k 64 * 128^k (>= 0), digit k of q is
// ((U >> 7k) & 127) - 64.  digit_bytes() returns the six digits of q as bytes d[0..5].
constexpr long long I8_OFFSET = 64ll * ((1ll << 42) - 1) / 127;  // sum_{k<6} 64 * 128^k
__device__ __forceinline__ void digit_bytes(long long q, uint32_t (&d)[I8_DIGITS]) {
  const unsigned long long U = (unsigned long long)(q + I8_OFFSET);
  const uint32_t lo = (uint32_t)U, hi = (uint32_t)(U >> 28);  // digits 0..3 from lo, 4..5 from bits 28..41
  d[0] = ((lo & 127u) - 64u) & 255u;
  d[1] = (((lo >> 7) & 127u) - 64u) & 255u;
  d[2] = (((lo >> 14) & 127u) - 64u) & 255u;
  d[3] = (((lo >> 21) & 127u) - 64u) & 255u;
  d[4] = ((hi & 127u) - 64u) & 255u;
  d[5] = (((hi >> 7) & 127u) - 64u) & 255u;
}
// tile = 32 columns x 128 rows; a thread digitises 4 consecutive rows of one column and stores one 32-bit word
// per plane; the planes go through shared memory so that the global writes (along i) are coalesced
__global__ void __launch_bounds__(256) digits_kernel(const double* __restrict__ X, int64_t N, int Ppad, int64_t Npad,
                                                     const double* __restrict__ qscale, int8_t* __restrict__ D8) {
  __shared__ __align__(4) int8_t sm[I8_DIGITS][32][132];
  const int p0 = blockIdx.x * 32;
  const int64_t i0 = (int64_t)blockIdx.y * 128;
  const int pl = threadIdx.x & 31;
  const int p = p0 + pl;
  const double sc = p < Ppad ? qscale[p] : 0.0;
  for (int i4 = threadIdx.x >> 5; i4 < 32; i4 += 8) {
    uint32_t word[I8_DIGITS] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t i = i0 + 4 * i4 + j;
      long long q = 0;
      if (i < N && p < Ppad) q = __double2ll_rn(X[i * Ppad + p] * sc);
      uint32_t d[I8_DIGITS];
      digit_bytes(q, d);
#pragma unroll
      for (int k = 0; k < I8_DIGITS; ++k) word[k] |= d[k] << (8 * j);
    }
#pragma unroll
    for (int k = 0; k < I8_DIGITS; ++k) *reinterpret_cast<uint32_t*>(&sm[k][pl][4 * i4]) = word[k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < I8_DIGITS * 32 * 32; e += 256) {
    const int w = e & 31, row = e >> 5, k = row >> 5, c = row & 31;
    const int64_t i = i0 + 4 * w;
    if (p0 + c < Ppad && i < Npad)
      *reinterpret_cast<uint32_t*>(D8 + ((int64_t)k * Ppad + p0 + c) * Npad + i) = *reinterpret_cast<const uint32_t*>(&sm[k][c][4 * w]);
  }
}
// multiplicities as int8 [nrep][Npad]; thread = 4 rows
__global__ void counts8_kernel(const uint32_t* __restrict__ counts, int64_t N, int64_t Npad, int64_t nrep,
                               int8_t* __restrict__ out, int* __restrict__ overflow) {
  const int64_t per_rep = Npad / 4;
  const int64_t total = nrep * per_rep;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / per_rep, i = (e - b * per_rep) * 4;
    uint32_t pk = 0;
    bool big = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t c = (i + j < N) ? counts[b * N + i + j] : 0u;
      big |= c > 127u;
      pk |= (c & 127u) << (8 * j);
    }
    if (big) *overflow = 1;
    *reinterpret_cast<uint32_t*>(out + b * Npad + i) = pk;
  }
}
// colsum[b][p] (+)= dscale_p * sum_k 128^k S[b][k*Ppad + p]
__global__ void digits_combine_kernel(const int32_t* __restrict__ S, int64_t nrep, int Ppad, const double* __restrict__ dscale,
                                      int accumulate, double* __restrict__ colsum) {
  const int64_t total = nrep * Ppad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / Ppad;
    const int p = (int)(e - b * Ppad);
    const int32_t* s = S + b * (int64_t)I8_DIGITS * Ppad + p;
    double v = 0.0;
#pragma unroll
    for (int k = I8_DIGITS - 1; k >= 0; --k) v = v * 128.0 + (double)s[(int64_t)k * Ppad];
    v *= dscale[p];
    colsum[e] = accumulate ? colsum[e] + v : v;
  }
}

// ---- the weighted Gram of a whole batch as ONE integer GEMM -----------------------------------------
// G_b[p][q] = sum_i c_bi (x~_ip x~_iq): the bootstrap multiplicities factor out of the second moments, so
// for all replicates of a batch the Gram tiles are  counts[nb x N] x Z[N x n_zcols],  Z = the pair-product
// columns of the model's tile set (diagonal tiles: upper triangle).  Z is digitised like x~ above
// (z 2^(40-e_p-e_q) rounded to an integer, six balanced base-128 digits, int8 planes), the GEMM runs on the
// tensor cores with exact int32 accumulation, and G = 2^(e_p+e_q-40) sum_k 128^k S_k.  Per element of Z the
// rounding is <= 2^-41 of the column's bound, random in sign: the sums are at least as accurate as fp64
// FMA accumulation over the same rows.  The fp64 gram_kernel remains for single fits, for models whose
// planes exceed the memory budget, and as the fallback for multiplicities above 127.
__global__ void zscale_kernel(int n_zcols, const int* __restrict__ zp, const int* __restrict__ zq,
                              const double* __restrict__ qscale, const double* __restrict__ dscale,
                              double* __restrict__ zqscale, double* __restrict__ zdscale) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_zcols) return;
  zqscale[c] = qscale[zp[c]] * qscale[zq[c]] * 9.094947017729282e-13;  // 2^-40 (all factors are powers of two)
  zdscale[c] = dscale[zp[c]] * dscale[zq[c]] * 1099511627776.0;         // 2^40
}
// rows [row0, row0 + ld) of the planes go to Z8[(k * n_zcols + col) * ld + (i - row0)]  (ld a multiple of 16)
__global__ void __launch_bounds__(256) zdigits_kernel(const double* __restrict__ X, int64_t N, int Ppad, int64_t row0,
                                                      int64_t ld, int n_zcols, const int* __restrict__ zp,
                                                      const int* __restrict__ zq, const double* __restrict__ zqscale,
                                                      int8_t* __restrict__ Z8) {
  __shared__ __align__(4) int8_t sm[I8_DIGITS][32][132];
  const int c0 = blockIdx.x * 32;
  const int64_t i0 = (int64_t)blockIdx.y * 128;  // local row
  const int cl = threadIdx.x & 31;
  const int col = c0 + cl;
  const bool col_ok = col < n_zcols;
  const int p = col_ok ? zp[col] : 0, q = col_ok ? zq[col] : 0;
  const double sc = col_ok ? zqscale[col] : 0.0;
  for (int i4 = threadIdx.x >> 5; i4 < 32; i4 += 8) {  // a thread digitises 4 consecutive rows: one word per plane
    uint32_t word[I8_DIGITS] = {0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t il = i0 + 4 * i4 + j, i = row0 + il;
      long long v = 0;
      if (il < ld && i < N && col_ok) v = __double2ll_rn(X[i * Ppad + p] * X[i * Ppad + q] * sc);
      uint32_t d[I8_DIGITS];
      digit_bytes(v, d);
#pragma unroll
      for (int k = 0; k < I8_DIGITS; ++k) word[k] |= d[k] << (8 * j);
    }
#pragma unroll
    for (int k = 0; k < I8_DIGITS; ++k) *reinterpret_cast<uint32_t*>(&sm[k][cl][4 * i4]) = word[k];
  }
  __syncthreads();
  for (int e = threadIdx.x; e < I8_DIGITS * 32 * 32; e += 256) {
    const int w = e & 31, row = e >> 5, k = row >> 5, c = row & 31;
    const int64_t i = i0 + 4 * w;
    if (c0 + c < n_zcols && i < ld)
      *reinterpret_cast<uint32_t*>(Z8 + ((int64_t)k * n_zcols + c0 + c) * ld + i) = *reinterpret_cast<const uint32_t*>(&sm[k][c][4 * w]);
  }
}
// G[b][zdst[c]] (+)= zdscale_c * sum_k 128^k S[b][k*n_zcols + c]   (and the mirrored entry of diagonal tiles)
__global__ void zcombine_kernel(const int32_t* __restrict__ S, int64_t nrep, int n_zcols, const double* __restrict__ zdscale,
                                const int* __restrict__ zdst, const int* __restrict__ zdst2, int accumulate,
                                int64_t g_stride, double* __restrict__ G) {
  const int64_t total = nrep * n_zcols;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / n_zcols;
    const int c = (int)(e - b * n_zcols);
    const int32_t* s = S + b * (int64_t)I8_DIGITS * n_zcols + c;
    double v = 0.0;
#pragma unroll
    for (int k = I8_DIGITS - 1; k >= 0; --k) v = v * 128.0 + (double)s[(int64_t)k * n_zcols];
    v *= zdscale[c];
    double* g = G + b * g_stride;
    const int d1 = zdst[c], d2 = zdst2[c];
    const double out = accumulate ? g[d1] + v : v;
    g[d1] = out;
    if (d2 >= 0) g[d2] = out;
  }
}

// Scores for the tensor-core sign vote:
//   B[i - i0][l*ldl + b] = fp16( c_bi * (x~_i . wf_b,l - sh_b,l) )      rows [i0, i0 + rc) of one chunk,
// ldl = replicates rounded up to 8: a thread's SG_RPT = 8 consecutive replicates are one 16-byte store.
// Only the SIGN of the resulting cross moments is used, and only where it exceeds a rigorous error bound
// (solver_core.h, phase 3), so the scores are computed in fp32 from an fp32 copy of x~: half the shared-
// memory traffic and staging of the fp64 version, twice the replicates per staged row tile.  The bound
// carries the fp32 term (k+4) 2^-24 sum_k |x_k w_k|.
// A group of nsl_pad adjacent lanes (power of two >= slots of the widest block) serves one (replicate
// lane, latent variable) pair: lane `sub` of the group owns slot `sub` of the block, keeps the weights
// of that slot for SG_RPT replicates in registers, walks the rows of the CTA's tiles reading the slot
// (8 floats) once for all of them, and the partial dot products are combined with a shuffle butterfly.
constexpr int SG_MAX_ROWS = 64, SG_RPT = 8, SG_THREADS = 256;
__global__ void make_float_kernel(const double* __restrict__ X, int64_t total, float* __restrict__ out) {
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = (float)X[e];
}
template <bool SINGLE_SLOT>  // every block fits one slot (nsl_pad == 1): no shuffle butterfly, idle lanes skip the rows
__global__ void __launch_bounds__(SG_THREADS) scoregen_kernel(const float* __restrict__ X,
                                                              const uint32_t* __restrict__ counts,
                                                              const double* __restrict__ wf,
                                                              const double* __restrict__ sh, int64_t N, int Ppad, int L,
                                                              const int* __restrict__ lv_off,
                                                              const int* __restrict__ lv_k, int nsl_pad, int SG_ROWS,
                                                              int64_t nrep, int64_t ldl, int64_t i0, int rc,
                                                              __half* __restrict__ B) {
  extern __shared__ __align__(16) float sg_smem[];
  float* xs = sg_smem;                                      // [SG_ROWS][Ppad]
  float* cs = xs + (size_t)SG_ROWS * Ppad;                  // [SG_ROWS][reps_per_cta] multiplicities
  int nbl = SG_THREADS / (L * nsl_pad);                     // replicate lanes per CTA (same rule on the host)
  if (SINGLE_SLOT && nbl >= 4) nbl = (SG_THREADS / 32 / ((L + 7) >> 3)) * 4;
  const int reps_per_cta = nbl * SG_RPT;
  const int64_t rep0 = (int64_t)blockIdx.y * reps_per_cta;
  // the staging stores of the multiplicities walk the rows (stride = one row of cs): XOR the group-of-four index
  // with the row so that they spread over the banks (power-of-two group counts only)
  const int groups4 = reps_per_cta / 4;
  const int swz = (groups4 & (groups4 - 1)) == 0 ? min(groups4, 8) - 1 : 0;
  // Thread -> (replicate lane bl, latent variable l, slot sub).  Single-slot blocks: a warp is 8 LVs x 4
  // replicate lanes, so the 32 LDS.128 of a row touch only 8 distinct (adjacent) slots.
  int sub, bl, l;
  bool active;
  if (SINGLE_SLOT && nbl >= 4) {
    const int lvg = (L + 7) >> 3, w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    l = (w % lvg) * 8 + (ln & 7);
    bl = (w / lvg) * 4 + (ln >> 3);
    sub = 0;
    active = l < L && bl < nbl;
    l = min(l, L - 1);
    bl = min(bl, nbl - 1);
  } else {
    const int item = threadIdx.x / nsl_pad;
    sub = threadIdx.x - item * nsl_pad;
    bl = min(item / L, nbl - 1);
    l = item % L;
    active = item < nbl * L;                                // (whole lane groups are active or not)
  }
  const bool has_slot = sub < ((lv_k[l] + SLOT - 1) >> 3);
  const int slot = (lv_off[l] >> 3) + (has_slot ? sub : 0);
  // a slot is two 16-byte chunks; slots 4 apart share shared-memory banks, so odd groups of four slots read
  // their chunks in the opposite order (the weights are permuted the same way: the dot product does not care)
  const int rot4 = ((slot >> 2) & 1) * 4;
  float w[SG_RPT][8], shv[SG_RPT];
  // the thread's SG_RPT replicates are adjacent in B (LV-major layout): one 16-byte store per row; replicates
  // past nrep (but inside the padded stride) get zeros
  const int64_t bb0 = rep0 + bl * SG_RPT;
  uint4* out = (active && sub == 0 && bb0 < ldl) ? reinterpret_cast<uint4*>(B + l * ldl + bb0) : nullptr;
#pragma unroll
  for (int j = 0; j < SG_RPT; ++j) {
    const int64_t bb = bb0 + j;
    const bool ok = bb < nrep;
    shv[j] = (ok && sub == 0) ? (float)sh[bb * L + l] : 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) w[j][k] = (ok && has_slot) ? (float)wf[bb * Ppad + slot * SLOT + (k ^ rot4)] : 0.f;
  }
  const int64_t ldb = ldl * L / SG_RPT;                      // row stride of B in 16-byte units
  const float* xcol = xs + slot * SLOT;
  // the block weights stay in registers while the CTA walks its share of the chunk's row tiles
  for (int row0 = blockIdx.x * SG_ROWS; row0 < rc; row0 += gridDim.x * SG_ROWS) {
    const int rows = min(SG_ROWS, rc - row0);
    __syncthreads();
    {
      const float4* src = reinterpret_cast<const float4*>(X + (i0 + row0) * Ppad);
      float4* dst = reinterpret_cast<float4*>(xs);
      const int n4 = rows * Ppad / 4;
      for (int e = threadIdx.x; e < n4; e += SG_THREADS) dst[e] = src[e];
    }
    for (int e = threadIdx.x; e < reps_per_cta * SG_ROWS; e += SG_THREADS) {
      const int eb = e / SG_ROWS, r = e - eb * SG_ROWS;     // consecutive threads: consecutive rows of one replicate
      const int64_t bb = rep0 + eb;
      float c = 0.f;
      if (r < rows && bb < nrep) c = counts ? (float)counts[bb * N + i0 + row0 + r] : 1.f;
      cs[r * reps_per_cta + (eb ^ (swz ? (r & swz) << 2 : 0))] = c;
    }
    __syncthreads();
    if (SINGLE_SLOT && !active) continue;  // (with lane groups everyone runs along: full-mask shuffles below)
    int64_t orow = (int64_t)row0 * ldb;
    for (int r = 0; r < rows; ++r, orow += ldb) {
      const float4 xa = *reinterpret_cast<const float4*>(xcol + (size_t)r * Ppad + rot4);
      const float4 xb = *reinterpret_cast<const float4*>(xcol + (size_t)r * Ppad + (4 - rot4));
      const int sw = swz ? (r & swz) << 2 : 0;              // undo the staging swizzle (groups of four replicates)
      const float4 ca = *reinterpret_cast<const float4*>(cs + r * reps_per_cta + ((bl * SG_RPT) ^ sw));
      const float4 cb = *reinterpret_cast<const float4*>(cs + r * reps_per_cta + ((bl * SG_RPT + 4) ^ sw));
      const float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      const float cj[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
      uint32_t pk[SG_RPT / 2];
#pragma unroll
      for (int j = 0; j < SG_RPT; j += 2) {
        float t0 = -shv[j], t1 = -shv[j + 1];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          t0 = fmaf(x[k], w[j][k], t0);
          t1 = fmaf(x[k], w[j + 1][k], t1);
        }
        if (!SINGLE_SLOT)
          for (int o = nsl_pad >> 1; o > 0; o >>= 1) {
            t0 += __shfl_xor_sync(0xffffffffu, t0, o);
            t1 += __shfl_xor_sync(0xffffffffu, t1, o);
          }
        const __half2 h2 = __floats2half2_rn(cj[j] * t0, cj[j + 1] * t1);
        pk[j / 2] = *reinterpret_cast<const uint32_t*>(&h2);
      }
      if (out) out[orow] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
    }
  }
}

// Stopping criterion of the non-metric path (weights.py:120), per replicate:
//   conv[b] = sum_l sum_i c_bi ( |y_old,il| - |y_new,il| )^2 ,   y = x~_i . coef_l - sh_l .
// (|a| - |b|)^2 = (a - b)^2 + 4ab [ab < 0]: the first part is a function of second moments and comes from
// num_step (conv_main); this pass adds  4 sum c y_old y_new  over the (row, LV) pairs whose score changes sign.
// Same thread mapping and tile walk as scoregen_kernel (lane groups of nsl_pad slots per (replicate lane, LV)).
// T = float: the scores are screened in fp32 from the fp32 copy of x~ (4 replicates per thread); any score
// within 8x its fp32 error bound of zero is recomputed in fp64 from X by the lane, so near the tolerance (where
// the score changes are tiny and every sign change is such a score) the result is the fp64 one; clear sign
// changes (both scores away from zero) only occur while the criterion is far above the tolerance and use the
// fp32 products (relative error 1e-5 of a number that is then >> tol).  T = double: everything in fp64 (N < 4096).
// Every CTA writes one partial per replicate; num_step_kernel adds the partials in a fixed order.
template <typename T> struct CvTraits;
template <> struct CvTraits<double> { static constexpr int RPT = 2; };
template <> struct CvTraits<float> { static constexpr int RPT = 4; };
template <typename T>
__global__ void __launch_bounds__(SG_THREADS, 2) conv_kernel(const T* __restrict__ Xs, const double* __restrict__ X,
                                                          const uint32_t* __restrict__ counts,
                                                          const double* __restrict__ coef_old,
                                                          const double* __restrict__ coef_new,
                                                          const double* __restrict__ sh_old,
                                                          const double* __restrict__ sh_new, const int* __restrict__ meta,
                                                          int64_t N, int Ppad, int L, const int* __restrict__ lv_off,
                                                          const int* __restrict__ lv_k, int nsl_pad, int ROWS,
                                                          int64_t nrep, double* __restrict__ conv_part) {
  constexpr int RPT = CvTraits<T>::RPT;
  constexpr bool F32 = sizeof(T) == 4;
  extern __shared__ __align__(16) unsigned char cv_smem_raw[];
  const int nbl = SG_THREADS / (L * nsl_pad);
  const int reps_per_cta = nbl * RPT;
  T* xs = reinterpret_cast<T*>(cv_smem_raw);                // [ROWS][Ppad]
  T* cs = xs + (size_t)ROWS * Ppad;                         // [ROWS][reps_per_cta]
  double* part = reinterpret_cast<double*>(cv_smem_raw + (((size_t)ROWS * (Ppad + reps_per_cta) * sizeof(T) + 15) & ~(size_t)15));
  const int64_t rep0 = (int64_t)blockIdx.y * reps_per_cta;
  const int item = threadIdx.x / nsl_pad, sub = threadIdx.x - item * nsl_pad;
  const int bl = min(item / L, nbl - 1), l = item % L;
  const bool active = item < nbl * L;
  const bool has_slot = sub < ((lv_k[l] + SLOT - 1) >> 3);
  const int slot = (lv_off[l] >> 3) + (has_slot ? sub : 0);
  // chunk order within the slot, rotated against shared-memory bank aliasing (double: 4 chunks of 2, float: 2 of 4)
  constexpr int CH = F32 ? 2 : 4, CW = 8 / CH;
  const int rot = F32 ? (slot >> 2) & 1 : (slot >> 1) & 3;
  T wo[RPT][8], wn[RPT][8], so[RPT], sn[RPT], nwo[RPT], nwn[RPT];
  double acc[RPT];
  bool live[RPT];
#pragma unroll
  for (int j = 0; j < RPT; ++j) {
    const int64_t bb = rep0 + bl * RPT + j;
    live[j] = bb < nrep && meta[bb * 4 + 1] == 0;  // finished replicates are skipped
    acc[j] = 0.0;
    so[j] = (live[j] && sub == 0) ? (T)sh_old[bb * L + l] : (T)0;
    sn[j] = (live[j] && sub == 0) ? (T)sh_new[bb * L + l] : (T)0;
    T no = 0, nn = 0;
#pragma unroll
    for (int ch = 0; ch < CH; ++ch)
#pragma unroll
      for (int e = 0; e < CW; ++e) {
        const int col = slot * SLOT + CW * ((ch + rot) % CH) + e;
        const bool ld = live[j] && has_slot;
        wo[j][CW * ch + e] = ld ? (T)coef_old[bb * Ppad + col] : (T)0;
        wn[j][CW * ch + e] = ld ? (T)coef_new[bb * Ppad + col] : (T)0;
        no += wo[j][CW * ch + e] * wo[j][CW * ch + e];
        nn += wn[j][CW * ch + e] * wn[j][CW * ch + e];
      }
    nwo[j] = sqrt(no); nwn[j] = sqrt(nn);
  }
  // fp32 screening threshold: 8 x the rounding bound (k+4) 2^-24 (|x_blk| |w_blk| + |sh|) of a score
  const T gam = (T)(8.0 * (8 * nsl_pad + 4) * 6.0e-8);
  for (int64_t row0 = (int64_t)blockIdx.x * ROWS; row0 < N; row0 += (int64_t)gridDim.x * ROWS) {
    const int rows = (int)min((int64_t)ROWS, N - row0);
    __syncthreads();
    {  // 16-byte copies (rows are multiples of 8 elements)
      const float4* src = reinterpret_cast<const float4*>(Xs + row0 * Ppad);
      float4* dst = reinterpret_cast<float4*>(xs);
      const int n4 = rows * Ppad * (int)sizeof(T) / 16;
      for (int e = threadIdx.x; e < n4; e += SG_THREADS) dst[e] = src[e];
    }
    for (int e = threadIdx.x; e < reps_per_cta * ROWS; e += SG_THREADS) {
      const int eb = e / ROWS, r = e - eb * ROWS;
      const int64_t bb = rep0 + eb;
      T c = 0;
      if (r < rows && bb < nrep) c = counts ? (T)counts[bb * N + row0 + r] : (T)1;
      cs[r * reps_per_cta + eb] = c;
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r) {
      T x[8];
      if constexpr (F32) {
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          const float4 v = *reinterpret_cast<const float4*>(xs + (size_t)r * Ppad + slot * SLOT + 4 * ((ch + rot) & 1));
          x[4 * ch] = v.x; x[4 * ch + 1] = v.y; x[4 * ch + 2] = v.z; x[4 * ch + 3] = v.w;
        }
      } else {
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const double2 v = *reinterpret_cast<const double2*>(xs + (size_t)r * Ppad + slot * SLOT + 2 * ((ch + rot) & 3));
          x[2 * ch] = v.x; x[2 * ch + 1] = v.y;
        }
      }
      T x2 = 0;
      if constexpr (F32) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x2 = fma(x[k], x[k], x2);
        x2 = sqrt(x2);
      }
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        T to = -so[j], tn = -sn[j];
#pragma unroll
        for (int k = 0; k < 8; ++k) { to = fma(x[k], wo[j][k], to); tn = fma(x[k], wn[j][k], tn); }
        T bo = F32 ? gam * (x2 * nwo[j] + fabs(so[j])) : (T)0, bn = F32 ? gam * (x2 * nwn[j] + fabs(sn[j])) : (T)0;
        for (int o = nsl_pad >> 1; o > 0; o >>= 1) {
          to += __shfl_xor_sync(0xffffffffu, to, o);
          tn += __shfl_xor_sync(0xffffffffu, tn, o);
          if constexpr (F32) {
            bo += __shfl_xor_sync(0xffffffffu, bo, o);
            bn += __shfl_xor_sync(0xffffffffu, bn, o);
          }
        }
        if (sub != 0 || !active || !live[j]) continue;
        const double c = (double)cs[r * reps_per_cta + bl * RPT + j];
        if constexpr (F32) {
          if (fabsf(to) > bo && fabsf(tn) > bn) {
            // both signs are certain.  A clear sign change needs |y_old - y_new| > 2 bound on this row, which only
            // happens while the criterion is orders of magnitude above the tolerance: fp32 products are enough there
            if (to * tn < 0.f) acc[j] = fma(4.0 * c * (double)to, (double)tn, acc[j]);
          } else {  // rare: a score within its fp32 error bound of zero -- exact scores of this (row, LV, replicate)
            const int64_t bb = rep0 + bl * RPT + j, i = row0 + r;
            double yo = -sh_old[bb * L + l], yn = -sh_new[bb * L + l];
            for (int q = lv_off[l]; q < lv_off[l] + lv_k[l]; ++q) {
              const double xv = X[i * Ppad + q];
              yo = fma(xv, coef_old[bb * Ppad + q], yo);
              yn = fma(xv, coef_new[bb * Ppad + q], yn);
            }
            if (yo * yn < 0.0) acc[j] = fma(4.0 * c * yo, yn, acc[j]);
          }
        } else {
          if (to * tn < 0.0) acc[j] = fma(4.0 * c * to, tn, acc[j]);
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < RPT; ++j) part[threadIdx.x * RPT + j] = (sub == 0 && active && live[j]) ? acc[j] : 0.0;
  __syncthreads();
  if (threadIdx.x < reps_per_cta) {  // fixed-order sum over the threads that served this replicate
    const int eb = threadIdx.x, ebl = eb / RPT, ej = eb - ebl * RPT;
    double s = 0.0;
    for (int ll = 0; ll < L; ++ll) s += part[((ebl * L + ll) * nsl_pad) * RPT + ej];
    if (rep0 + eb < nrep) conv_part[(rep0 + eb) * gridDim.x + blockIdx.x] = s;
  }
}

struct NumBatch {
  ModelView M;
  const double* G; int64_t g_stride;
  const double* colsum;
  double N;
  int scheme; double tol; int max_iter;
  const double* conv_part; int n_conv_part;
  double* conv_main;  // [nrep] second-moment part of the criterion, written by num_step
  double* ws;
  double *a, *coef_old, *coef_new, *shift_old, *shift_new;
  int* meta;
  int* n_done;
  double* out_rows; int64_t out_stride;
  double *weights, *loadings, *r2, *paths, *total, *crossloadings, *score_coef, *score_shift;
  int *iters, *status;
};

__global__ void __launch_bounds__(128) num_step_kernel(const NumBatch b) {
  extern __shared__ __align__(16) double solver_smem_num[];
  const int64_t rep = blockIdx.x;
  if (b.meta[rep * 4 + 1]) return;
  NumStepArgs A;
  A.M = b.M;
  A.G = b.G + rep * b.g_stride;
  A.colsum = b.colsum + rep * b.M.Ppad;
  A.N = b.N; A.scheme = b.scheme; A.tol = b.tol; A.max_iter = b.max_iter;
  double conv = b.conv_main[rep];
  for (int k = 0; k < b.n_conv_part; ++k) conv += b.conv_part[rep * b.n_conv_part + k];
  A.conv_in = conv;
  A.conv_main = b.conv_main + rep;
  A.ws = b.ws + rep * (int64_t)b.M.ws_doubles;
  A.a = b.a + rep * b.M.Ppad;
  A.meta = b.meta + rep * 4;
  A.coef_old = b.coef_old + rep * b.M.Ppad; A.coef_new = b.coef_new + rep * b.M.Ppad;
  A.shift_old = b.shift_old + rep * b.M.L; A.shift_new = b.shift_new + rep * b.M.L;
  A.out_row = b.out_rows ? b.out_rows + rep * b.out_stride : nullptr;
  A.weights = b.weights; A.loadings = b.loadings; A.r2 = b.r2; A.paths = b.paths; A.total = b.total;
  A.crossloadings = b.crossloadings; A.score_coef = b.score_coef; A.score_shift = b.score_shift;
  A.iters = b.iters + rep; A.status = b.status + rep;
  num_step(A, solver_smem_num);
  __syncthreads();
  if (threadIdx.x == 0 && A.meta[1]) atomicAdd(b.n_done, 1);
}

// ------------------------------------------------------------------------------------------------
// weighted Gram kernel
// ------------------------------------------------------------------------------------------------
constexpr int GRAM_WARPS = 8;    // consumer warps per CTA, each = one (replicate, tile group) item
constexpr int GRAM_THREADS = GRAM_WARPS * 32;  // 2 warps per SM sub-partition: up to 255 registers per thread
constexpr int GRAM_MAX_STAGES = 8;

struct GramParams {
  const double* X;          // [N][Ppad]
  const uint32_t* counts;   // [nrep][N] or null (every row once)
  int64_t N;
  int Ppad, n_tiles, n_tg;
  const int *tile_sa, *tile_sb, *lane_tile;
  int64_t n_items;          // nrep * n_tg
  int n_chunks;
  int64_t chunk_rows;       // multiple of RT
  int RT, stages;
  double* G;                // [nrep][n_chunks][n_tiles*64]
  // cross-moment mode (template CROSS): tiles are (slot sa, LV group g) in natural order, the
  // column operand is the row's LV scores x~_i . wf_l times the multiplicity (per-warp scratch)
  int L, ng;
  const int *lv_off, *lv_k;
  const double* wf;         // [nrep][Ppad] final weights of every replicate
  const int* rep_map;       // optional: item / n_tg -> replicate (exact redo of selected replicates)
};

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}

template <bool CROSS>
__global__ void __launch_bounds__(GRAM_THREADS, 1) gram_kernel(const GramParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[GRAM_MAX_STAGES];
  double* tiles = reinterpret_cast<double*>(smem_raw);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_groups = (p.n_items + GRAM_WARPS - 1) / GRAM_WARPS;
  const int chunk = (int)(blockIdx.x / n_groups);
  const int64_t group = blockIdx.x - (int64_t)chunk * n_groups;
  const int64_t item0 = group * GRAM_WARPS;
  const int n_active = (int)min((int64_t)GRAM_WARPS, p.n_items - item0);
  const int64_t r0 = (int64_t)chunk * p.chunk_rows, r1 = min(r0 + p.chunk_rows, p.N);
  const int n_rt = (int)((r1 - r0 + p.RT - 1) / p.RT);
  const size_t stage_doubles = (size_t)p.RT * p.Ppad;

  // ---- feeding the ring -------------------------------------------------------------------------
  // No producer warp (a ninth warp would put three warps on one SM sub-partition and cap every thread at
  // 168 registers; the 8x8 fp64 accumulator tile alone needs 128).  Thread 0 issues the first `stages`
  // row tiles; after that the LAST consumer warp to finish with a stage (shared-memory counter) refills it
  // through the TMA engine at once, so a tile is always requested stages-1 tile times ahead of its use
  // no matter how the warps drift apart.
  __shared__ int stage_done[GRAM_MAX_STAGES];
  auto issue_tile = [&](int tn) {
    const int st = tn % p.stages;
    const int64_t row = r0 + (int64_t)tn * p.RT;
    const uint32_t rows = (uint32_t)min((int64_t)p.RT, r1 - row);
    const uint32_t bytes = rows * (uint32_t)p.Ppad * 8u;
    mbar_arrive_expect_tx(&full_bar[st], bytes);
    bulk_g2s(tiles + (size_t)st * stage_doubles, p.X + row * p.Ppad, bytes, &full_bar[st]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      stage_done[s] = 0;
    }
    mbar_fence_init();
    for (int tn = 0; tn < n_rt && tn < p.stages; ++tn) issue_tile(tn);
  }
  __syncthreads();
  if (warp >= n_active) return;

  // ---- consumer warp: one (replicate, tile group); lane = one 8x8 tile ------------------------
  // items are tile-group major: the warps of a CTA work on the same tile group for 8 replicates, so a CTA
  // is (except at a group boundary) all "diagonal" or all "generic" warps -- see below
  const int64_t item = item0 + warp;
  const int64_t nrep_pos = p.n_items / p.n_tg;
  const int tg = (int)(item / nrep_pos);
  const int64_t rep_pos = item - (int64_t)tg * nrep_pos;
  const int64_t rep = p.rep_map ? (int64_t)p.rep_map[rep_pos] : rep_pos;
  int tile, sa, sb;
  if constexpr (CROSS) {
    // LV-group-major order: a warp covers (almost always) ONE group of 8 LVs and 32 row slots, so it
    // needs only that group's scores
    tile = tg * 32 + lane;
    if (tile >= p.n_tiles) tile = -1;
    const int ns = p.Ppad / SLOT;
    sb = tile >= 0 ? tile / ns : 0;          // LV group: "slot" sb of the score scratch row
    sa = tile >= 0 ? tile - sb * ns : 0;
  } else {
    tile = p.lane_tile[tg * 32 + lane];
    sa = tile >= 0 ? p.tile_sa[tile] : 0;
    sb = tile >= 0 ? p.tile_sb[tile] : 0;
  }
  const bool tile_ok = tile >= 0;
  // A tile group that holds only diagonal tiles (sa == sb; the model builder packs them together) needs one
  // operand per row and, by symmetry, 36 of the 64 products.
  const bool diag = !CROSS && __all_sync(0xffffffffu, !tile_ok || sa == sb);
  // 16-byte chunks of a slot are read in a lane-dependent rotated order so that the 32 LDS.128 of
  // a warp spread over all bank quads (slot stride 64 B would otherwise be a 16-way conflict).
  // This holds for the row operand too: a sparse tile group holds ~3 tiles per row slot, i.e. ~11
  // distinct row slots per warp (ncu: 8.0 wavefronts per unrotated xa LDS.128 vs 4.33 rotated; the
  // kernel is bound by shared-memory wavefronts, 94 % L1/TEX throughput, before the fp64 pipe).
  const int rot_a = (sa >> 1) & 3, rot_b = CROSS ? 0 : ((sb >> 1) & 3);
  int off_a[4], off_b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    off_a[k] = sa * SLOT + 2 * ((k + rot_a) & 3);
    off_b[k] = sb * SLOT + 2 * ((k + rot_b) & 3);
  }
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

  const uint32_t* cnt_row = p.counts ? p.counts + rep * p.N : nullptr;
  auto load_counts = [&](int t_load) -> uint32_t {
    const int64_t row = r0 + (int64_t)t_load * p.RT;
    const int rows = (int)min((int64_t)p.RT, r1 - row);
    if (lane >= rows) return 0u;
    return cnt_row ? __ldg(cnt_row + row + lane) : 1u;
  };
  // shared-memory byte addresses of the lane's 2 x 4 16-byte operand chunks within a row
  uint32_t boff_a[4], boff_b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    boff_a[k] = (uint32_t)off_a[k] * 8u;
    boff_b[k] = (uint32_t)off_b[k] * 8u;
  }
  auto lds128 = [](uint32_t addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
  };
  auto load_row = [&](uint32_t row_addr, uint32_t b_addr, double (&xa)[8], double (&xb)[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lds128(row_addr + boff_a[k], xa[2 * k], xa[2 * k + 1]);
      lds128(b_addr + boff_b[k], xb[2 * k], xb[2 * k + 1]);
    }
  };
  // The scaling and the 64 FMAs of a row are emitted as volatile asm so that the compiler keeps
  // them AFTER the (volatile) shared-memory loads of the NEXT row in program order: without this
  // the loads get sunk below the FMA block to save registers and the software pipeline is lost
  // (measured: 6.5 % of all issue slots stalled on the first DMUL of every row).
  // `scale` is a compile-time tag: rows of multiplicity 1 (58 % of the non-zero rows of a resample) are
  // listed first and skip the 8 multiplications
  auto accumulate = [&](auto scale, double (&xa)[8], double (&xb)[8], double c) {
    if constexpr (!CROSS && decltype(scale)::value)  // (the score scratch is already multiplied by the multiplicity)
    asm volatile(
        "mul.f64 %0, %0, %8;\n\tmul.f64 %1, %1, %8;\n\tmul.f64 %2, %2, %8;\n\tmul.f64 %3, %3, %8;\n\t"
        "mul.f64 %4, %4, %8;\n\tmul.f64 %5, %5, %8;\n\tmul.f64 %6, %6, %8;\n\tmul.f64 %7, %7, %8;"
        : "+d"(xb[0]), "+d"(xb[1]), "+d"(xb[2]), "+d"(xb[3]), "+d"(xb[4]), "+d"(xb[5]), "+d"(xb[6]), "+d"(xb[7])
        : "d"(c));
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "fma.rn.f64 %0, %8, %9, %0;\n\tfma.rn.f64 %1, %8, %10, %1;\n\tfma.rn.f64 %2, %8, %11, %2;\n\t"
          "fma.rn.f64 %3, %8, %12, %3;\n\tfma.rn.f64 %4, %8, %13, %4;\n\tfma.rn.f64 %5, %8, %14, %5;\n\t"
          "fma.rn.f64 %6, %8, %15, %6;\n\tfma.rn.f64 %7, %8, %16, %7;"
          : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3]), "+d"(acc[i][4]), "+d"(acc[i][5]),
            "+d"(acc[i][6]), "+d"(acc[i][7])
          : "d"(xa[i]), "d"(xb[0]), "d"(xb[1]), "d"(xb[2]), "d"(xb[3]), "d"(xb[4]), "d"(xb[5]), "d"(xb[6]),
            "d"(xb[7]));
  };
  auto accumulate_diag = [&](auto scale, double (&xa)[8], double (&xs)[8], double c) {  // xs = c * xa, upper triangle only
    if constexpr (!decltype(scale)::value) {
#pragma unroll
      for (int k = 0; k < 8; ++k) xs[k] = xa[k];
    } else
    asm volatile(
        "mul.f64 %0, %8, %16;\n\tmul.f64 %1, %9, %16;\n\tmul.f64 %2, %10, %16;\n\tmul.f64 %3, %11, %16;\n\t"
        "mul.f64 %4, %12, %16;\n\tmul.f64 %5, %13, %16;\n\tmul.f64 %6, %14, %16;\n\tmul.f64 %7, %15, %16;"
        : "=d"(xs[0]), "=d"(xs[1]), "=d"(xs[2]), "=d"(xs[3]), "=d"(xs[4]), "=d"(xs[5]), "=d"(xs[6]), "=d"(xs[7])
        : "d"(xa[0]), "d"(xa[1]), "d"(xa[2]), "d"(xa[3]), "d"(xa[4]), "d"(xa[5]), "d"(xa[6]), "d"(xa[7]), "d"(c));
    asm volatile("fma.rn.f64 %0, %8, %9, %0;\n\tfma.rn.f64 %1, %8, %10, %1;\n\tfma.rn.f64 %2, %8, %11, %2;\n\tfma.rn.f64 %3, %8, %12, %3;\n\tfma.rn.f64 %4, %8, %13, %4;\n\tfma.rn.f64 %5, %8, %14, %5;\n\tfma.rn.f64 %6, %8, %15, %6;\n\tfma.rn.f64 %7, %8, %16, %7;"
                 : "+d"(acc[0][0]), "+d"(acc[0][1]), "+d"(acc[0][2]), "+d"(acc[0][3]), "+d"(acc[0][4]), "+d"(acc[0][5]), "+d"(acc[0][6]), "+d"(acc[0][7])
                 : "d"(xa[0]), "d"(xs[0]), "d"(xs[1]), "d"(xs[2]), "d"(xs[3]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %7, %8, %0;\n\tfma.rn.f64 %1, %7, %9, %1;\n\tfma.rn.f64 %2, %7, %10, %2;\n\tfma.rn.f64 %3, %7, %11, %3;\n\tfma.rn.f64 %4, %7, %12, %4;\n\tfma.rn.f64 %5, %7, %13, %5;\n\tfma.rn.f64 %6, %7, %14, %6;"
                 : "+d"(acc[1][1]), "+d"(acc[1][2]), "+d"(acc[1][3]), "+d"(acc[1][4]), "+d"(acc[1][5]), "+d"(acc[1][6]), "+d"(acc[1][7])
                 : "d"(xa[1]), "d"(xs[1]), "d"(xs[2]), "d"(xs[3]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %6, %7, %0;\n\tfma.rn.f64 %1, %6, %8, %1;\n\tfma.rn.f64 %2, %6, %9, %2;\n\tfma.rn.f64 %3, %6, %10, %3;\n\tfma.rn.f64 %4, %6, %11, %4;\n\tfma.rn.f64 %5, %6, %12, %5;"
                 : "+d"(acc[2][2]), "+d"(acc[2][3]), "+d"(acc[2][4]), "+d"(acc[2][5]), "+d"(acc[2][6]), "+d"(acc[2][7])
                 : "d"(xa[2]), "d"(xs[2]), "d"(xs[3]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %5, %6, %0;\n\tfma.rn.f64 %1, %5, %7, %1;\n\tfma.rn.f64 %2, %5, %8, %2;\n\tfma.rn.f64 %3, %5, %9, %3;\n\tfma.rn.f64 %4, %5, %10, %4;"
                 : "+d"(acc[3][3]), "+d"(acc[3][4]), "+d"(acc[3][5]), "+d"(acc[3][6]), "+d"(acc[3][7])
                 : "d"(xa[3]), "d"(xs[3]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %4, %5, %0;\n\tfma.rn.f64 %1, %4, %6, %1;\n\tfma.rn.f64 %2, %4, %7, %2;\n\tfma.rn.f64 %3, %4, %8, %3;"
                 : "+d"(acc[4][4]), "+d"(acc[4][5]), "+d"(acc[4][6]), "+d"(acc[4][7])
                 : "d"(xa[4]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %3, %4, %0;\n\tfma.rn.f64 %1, %3, %5, %1;\n\tfma.rn.f64 %2, %3, %6, %2;"
                 : "+d"(acc[5][5]), "+d"(acc[5][6]), "+d"(acc[5][7])
                 : "d"(xa[5]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %2, %3, %0;\n\tfma.rn.f64 %1, %2, %4, %1;"
                 : "+d"(acc[6][6]), "+d"(acc[6][7])
                 : "d"(xa[6]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %1, %2, %0;"
                 : "+d"(acc[7][7])
                 : "d"(xa[7]), "d"(xs[7]));
  };
  auto load_row_diag = [&](uint32_t row_addr, double (&xa)[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) lds128(row_addr + boff_a[k], xa[2 * k], xa[2 * k + 1]);
  };
  // per-warp list of the tile's non-zero rows: {row byte offset in the stage, multiplicity as fp64},
  // built once per tile by all lanes, so that the row loop is a plain counted loop
  __shared__ __align__(16) double2 row_list[GRAM_WARPS][40];  // 32 rows + 8 zero-multiplicity pads
  double2* my_list = row_list[warp];
  const uint32_t list_addr = smem_u32(my_list);
  const uint32_t tiles_addr = smem_u32(tiles);
  const uint32_t row_bytes = (uint32_t)p.Ppad * 8u;
  auto load_entry = [&](int k, uint32_t& off, double& c) {
    double o;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o), "=d"(c) : "r"(list_addr + 16u * (uint32_t)k));
    off = (uint32_t)__double_as_longlong(o);
  };
  // cross mode: per-warp scratch of (RT + 8) rows x Lpad scores behind the ring
  const int Lpad = p.ng * SLOT;
  double* my_scores = tiles + (size_t)p.stages * stage_doubles + (size_t)warp * (p.RT + 8) * Lpad;
  const uint32_t sc_addr = smem_u32(my_scores);
  const uint32_t sc_row_bytes = (uint32_t)Lpad * 8u;
  const double* wf_rep = CROSS ? p.wf + rep * p.Ppad : nullptr;
  if constexpr (CROSS) {
    for (int e = lane; e < (p.RT + 8) * Lpad; e += 32) my_scores[e] = 0.0;
    __syncwarp();
  }
  uint32_t cnt_next = load_counts(0);

  for (int t = 0; t < n_rt; ++t) {
    const int s = t % p.stages;
    const uint32_t use = (uint32_t)(t / p.stages);
    const uint32_t cnt = cnt_next;
    if (t + 1 < n_rt) cnt_next = load_counts(t + 1);  // prefetch: hides the global-load latency
    // rows of multiplicity 1 first, then the others
    const uint32_t mask1 = __ballot_sync(0xffffffffu, cnt == 1), mask2 = __ballot_sync(0xffffffffu, cnt > 1);
    const int n_one = __popc(mask1), n_nz = n_one + __popc(mask2);
    if (cnt != 0) {
      const uint32_t below = (1u << lane) - 1u;
      const int pos = (cnt == 1) ? __popc(mask1 & below) : n_one + __popc(mask2 & below);
      my_list[pos] = make_double2(__longlong_as_double((long long)((uint32_t)lane * row_bytes)), (double)cnt);
    }
    const int n_pairs_one = CROSS ? 0 : (n_one & ~1);  // an odd last multiplicity-1 row takes the scaled path
    // pads: multiplicity 0 on row 0 of the stage, so the pipelined loop below needs no branches
    if (lane < 8) my_list[n_nz + lane] = make_double2(__longlong_as_double(0ll), 0.0);
    __syncwarp();
    mbar_wait(&full_bar[s], use & 1);
    const uint32_t base = tiles_addr + (uint32_t)s * (uint32_t)(stage_doubles * 8);
    uint32_t release_dep = 0;
    if constexpr (CROSS) {
      // scores of the tile's non-zero rows for the LVs this warp's tiles touch, one (row, LV) pair
      // per lane:  scratch[k][l] = c_k * sum_{q in block l} x~[row_k][q] wf[q]   (pad rows: c = 0)
      if (n_nz > 0) {
        const int ns = p.Ppad / SLOT;
        const int lv_lo = ((tg * 32) / ns) * SLOT;
        const int lv_hi = min(p.L, (min(p.n_tiles - 1, tg * 32 + 31) / ns) * SLOT + SLOT);
        const int nlw = lv_hi - lv_lo;
        for (int e = lane; e < nlw * (n_nz + 2); e += 32) {
          const int k = e / nlw, lv = lv_lo + (e - k * nlw);
          const int slot0 = p.lv_off[lv] >> 3, nsl = (p.lv_k[lv] + SLOT - 1) >> 3;
          const int rot = (slot0 >> 1) & 3;
          uint32_t o;
          double cc, sc = 0.0;
          load_entry(k, o, cc);
          for (int sl = 0; sl < nsl; ++sl)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const int col = (slot0 + sl) * SLOT + 2 * ((ch + rot) & 3);
              const double2 w2 = __ldg(reinterpret_cast<const double2*>(wf_rep + col));
              double x0, x1;
              lds128(base + o + (uint32_t)col * 8u, x0, x1);
              sc = fma(x0, w2.x, sc);
              sc = fma(x1, w2.y, sc);
            }
          my_scores[(size_t)k * Lpad + lv] = cc * sc;
        }
      }
      __syncwarp();
    }
    if (n_nz > 0 && diag) {
      // diagonal tile group: same software pipeline, one operand per row, 8 DMUL + 36 DFMA per row
      double xa0[8], xa1[8], xs[8], c0, c1, ce0, ce1;
      uint32_t oe0, oe1;
      load_entry(0, oe0, ce0);
      load_entry(1, oe1, ce1);
      load_row_diag(base + oe0, xa0);
      c0 = ce0;
      int k = 0;
      for (; k < n_pairs_one; k += 2) {
        load_row_diag(base + oe1, xa1);
        load_entry(k + 2, oe0, ce0);
        accumulate_diag(std::false_type{}, xa0, xs, 1.0);
        load_row_diag(base + oe0, xa0);
        load_entry(k + 3, oe1, ce1);
        accumulate_diag(std::false_type{}, xa1, xs, 1.0);
      }
      c0 = ce0;
      for (; k < n_nz; k += 2) {
        load_row_diag(base + oe1, xa1);
        c1 = ce1;
        load_entry(k + 2, oe0, ce0);
        accumulate_diag(std::true_type{}, xa0, xs, c0);
        load_row_diag(base + oe0, xa0);
        c0 = ce0;
        load_entry(k + 3, oe1, ce1);
        accumulate_diag(std::true_type{}, xa1, xs, c1);
      }
      asm volatile("{\n.reg .b32 lo, hi;\nmov.b64 {lo, hi}, %1;\nand.b32 %0, lo, 0;\n}" : "=r"(release_dep) : "d"(xa0[7]));
    } else if (n_nz > 0) {
      // Software pipeline over the non-zero rows, two rows per trip, straight-line body: the
      // operands of row k+1 are in flight (LDS) while the 64 FMAs of row k issue, and the list
      // entries of rows k+2 / k+3 are already in registers.  An odd row count runs one padded row
      // with multiplicity 0 (adds exact zeros; ~2.5 % extra FMAs, no branch in the body).
      double xa0[8], xb0[8], xa1[8], xb1[8], c0, c1, ce0, ce1;
      uint32_t oe0, oe1;
      // column operand: the same X row (Gram) or the row's scratch scores (cross)
      uint32_t bsrc = sc_addr;
      load_entry(0, oe0, ce0);
      load_entry(1, oe1, ce1);
      load_row(base + oe0, CROSS ? bsrc : base + oe0, xa0, xb0);
      c0 = ce0;
      int k = 0;
      for (; k < n_pairs_one; k += 2) {  // multiplicity 1: no scaling (Gram mode only)
        load_row(base + oe1, base + oe1, xa1, xb1);
        load_entry(k + 2, oe0, ce0);
        accumulate(std::false_type{}, xa0, xb0, 1.0);
        load_row(base + oe0, base + oe0, xa0, xb0);
        load_entry(k + 3, oe1, ce1);
        accumulate(std::false_type{}, xa1, xb1, 1.0);
      }
      c0 = ce0;
      for (; k < n_nz; k += 2) {
        load_row(base + oe1, CROSS ? bsrc + sc_row_bytes : base + oe1, xa1, xb1);
        c1 = ce1;
        load_entry(k + 2, oe0, ce0);
        accumulate(std::true_type{}, xa0, xb0, c0);
        bsrc += 2 * sc_row_bytes;
        load_row(base + oe0, CROSS ? bsrc : base + oe0, xa0, xb0);
        c0 = ce0;
        load_entry(k + 3, oe1, ce1);
        accumulate(std::true_type{}, xa1, xb1, c1);
      }
      // The loop prefetches one row set past the end (a pad row of this stage).  Make the stage release
      // below depend on that last load, so no shared-memory read of the stage is still in flight when
      // the producer's next bulk copy may overwrite it.
      asm volatile("{\n.reg .b32 lo, hi;\nmov.b64 {lo, hi}, %1;\nand.b32 %0, lo, 0;\n}" : "=r"(release_dep) : "d"(xb0[7]));
    }
    __syncwarp();
    if (lane == 0) {
      // (release_dep == 0, but it makes this release depend on the warp's last shared-memory load)
      __threadfence_block();
      const int prev = atomicAdd(&stage_done[s] + release_dep, 1);
      if (prev == n_active - 1) {  // every consumer is done with this fill: refill the stage
        stage_done[s] = 0;
        if (t + p.stages < n_rt) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue_tile(t + p.stages);
        }
      }
    }
  }

  // ---- write the partial tile (undo the chunk rotation) and the column sums --------------------
  const size_t slab = (size_t)rep * p.n_chunks + chunk;
  if (tile_ok) {
    const int store_tile = CROSS ? sa * p.ng + sb : tile;
    double* g = p.G + (slab * p.n_tiles + store_tile) * TILE;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ra = 2 * (((i >> 1) + rot_a) & 3) + (i & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cb = 2 * (((j >> 1) + rot_b) & 3) + (j & 1);
        g[ra * SLOT + cb] = (diag && j < i) ? acc[j][i] : acc[i][j];  // diagonal groups hold the upper triangle
      }
    }
  }
}

// Weighted column sums colsum[b][p] = sum_i c_bi x~_ip  (a skinny fp64 GEMM, counts x X~).
// CTA = 32 replicates x 256 columns over one row chunk; the X row tile and the (fp64-converted)
// multiplicities are staged in shared memory, thread = 8 replicates x 4 columns in registers
// (6 LDS.128 per 32 FMAs), two CTAs per SM overlap staging and arithmetic.
constexpr int CS_REPS = 32, CS_COLS = 256, CS_ROWS = 32;
__global__ void __launch_bounds__(256, 2) colsum_kernel(const double* __restrict__ X, const uint32_t* __restrict__ counts,
                                                        int64_t N, int Ppad, int64_t nrep, int n_chunks,
                                                        int64_t chunk_rows, double* __restrict__ out) {
  extern __shared__ __align__(16) double cs_smem[];
  double* xs = cs_smem;                       // [CS_ROWS][CS_COLS]
  double* cw = xs + CS_ROWS * CS_COLS;        // [CS_ROWS][CS_REPS]
  const int col0 = blockIdx.x * CS_COLS;
  const int64_t rep0 = (int64_t)blockIdx.y * CS_REPS;
  const int chunk = blockIdx.z;
  const int64_t r0 = (int64_t)chunk * chunk_rows, r1 = min(r0 + chunk_rows, N);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = warp >> 1;                   // replicate group: replicates 8*rg .. 8*rg+7
  const int cg = (warp & 1) * 32 + lane;      // column group: columns 4*cg .. 4*cg+3
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int64_t row = r0; row < r1; row += CS_ROWS) {
    const int rows = (int)min((int64_t)CS_ROWS, r1 - row);
    __syncthreads();
    for (int e = threadIdx.x; e < CS_ROWS * CS_COLS; e += 256) {
      const int r = e / CS_COLS, c = e - r * CS_COLS;
      xs[e] = (r < rows && col0 + c < Ppad) ? X[(row + r) * Ppad + col0 + c] : 0.0;
    }
    for (int e = threadIdx.x; e < CS_ROWS * CS_REPS; e += 256) {
      const int b = e / CS_ROWS, r = e - b * CS_ROWS;  // consecutive threads: consecutive rows of one replicate
      double v = 0.0;
      if (r < rows && rep0 + b < nrep) v = counts ? (double)counts[(rep0 + b) * N + row + r] : 1.0;
      cw[r * CS_REPS + b] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < CS_ROWS; ++r) {
      const double2 x01 = *reinterpret_cast<const double2*>(&xs[r * CS_COLS + 4 * cg]);
      const double2 x23 = *reinterpret_cast<const double2*>(&xs[r * CS_COLS + 4 * cg + 2]);
      double c[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double2 c2 = *reinterpret_cast<const double2*>(&cw[r * CS_REPS + 8 * rg + 2 * k]);
        c[2 * k] = c2.x; c[2 * k + 1] = c2.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fma(c[i], x01.x, acc[i][0]); acc[i][1] = fma(c[i], x01.y, acc[i][1]);
        acc[i][2] = fma(c[i], x23.x, acc[i][2]); acc[i][3] = fma(c[i], x23.y, acc[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t b = rep0 + 8 * rg + i;
    if (b >= nrep) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + 4 * cg + j;
      if (c < Ppad) out[(b * n_chunks + chunk) * Ppad + c] = acc[i][j];
    }
  }
}

// sum of the per-chunk partials in chunk order (deterministic)
__global__ void reduce_chunks_kernel(const double* __restrict__ part, int64_t nrep, int n_chunks, int64_t per_rep,
                                     double* __restrict__ out) {
  const int64_t total = nrep * per_rep;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / per_rep, k = e - b * per_rep;
    const double* src = part + (b * n_chunks) * per_rep + k;
    double s = 0.0;
    for (int c = 0; c < n_chunks; ++c) s += src[(int64_t)c * per_rep];
    out[e] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// per-replicate solver kernel (one CTA per replicate)
// ------------------------------------------------------------------------------------------------
struct SolveBatch {
  ModelView M;
  const double* G; int64_t g_stride;
  const double* colsum; int64_t cs_stride;
  const double* mu;
  double N;
  int scheme; double tol; int max_iter;
  double* ws;
  int phase;                             // see SolveArgs::phase
  double* wf;                            // [nrep][Ppad] (sparse tile sets)
  const double* cross; int64_t cross_stride;
  const float* fast_cross; const double* inv_sd; int64_t fast_nb;  // phase 3
  double* sh;                            // [nrep][L] (phase 1 output)
  const int* rep_map;                    // optional: block -> replicate
  double* out_rows; int64_t out_stride;  // may be null
  double *weights, *loadings, *r2, *paths, *total, *crossloadings, *score_coef, *score_shift;  // single fit
  int *iters, *status;
};

constexpr int SOLVE_THREADS = 128;

__global__ void __launch_bounds__(SOLVE_THREADS) solve_kernel(const SolveBatch b) {
  extern __shared__ __align__(16) double solver_smem[];
  const int64_t rep = b.rep_map ? (int64_t)b.rep_map[blockIdx.x] : (int64_t)blockIdx.x;
  SolveArgs A;
  A.M = b.M;
  A.G = b.G + rep * b.g_stride;
  A.colsum = b.colsum + rep * b.cs_stride;
  A.mu = b.mu;
  A.N = b.N;
  A.scheme = b.scheme;
  A.tol = b.tol;
  A.max_iter = b.max_iter;
  A.phase = b.phase;
  A.wf_out = b.wf ? b.wf + rep * b.M.Ppad : nullptr;
  A.cross = b.cross ? b.cross + rep * b.cross_stride : nullptr;
  A.fast_cross = b.fast_cross; A.inv_sd = b.inv_sd; A.fast_nb = b.fast_nb; A.fast_b = rep;
  A.sh_out = b.sh ? b.sh + rep * b.M.L : nullptr;
  A.ws = b.ws + rep * (int64_t)b.M.ws_doubles;
  A.out_row = b.out_rows ? b.out_rows + rep * b.out_stride : nullptr;
  A.weights = b.weights; A.loadings = b.loadings; A.r2 = b.r2; A.paths = b.paths; A.total = b.total;
  A.crossloadings = b.crossloadings; A.score_coef = b.score_coef; A.score_shift = b.score_shift;
  A.iters = b.iters + rep;
  A.status = b.status + rep;
  solve_replicate(A, solver_smem);
}

// scores[i][l] = sum_{c in block l} x~[i][c] coef[c] - shift[l]     (weights.py:60, 65-68)
__global__ void scores_kernel(const double* __restrict__ X, int64_t N, int Ppad, int L, const int* __restrict__ lv_off,
                              const int* __restrict__ lv_k, const double* __restrict__ coef,
                              const double* __restrict__ shift, double* __restrict__ scores) {
  const int64_t total = N * L;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = e / L;
    const int l = (int)(e - i * L);
    const int o = lv_off[l], k = lv_k[l];
    const double* x = X + i * Ppad + o;
    const double* cf = coef + o;
    double s = 0.0;
    for (int c = 0; c < k; ++c) s = fma(x[c], cf[c], s);
    scores[e] = s - shift[l];
  }
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int plspm_version(void) { return 100; }
const char* plspm_last_error(void) { return g_err.c_str(); }

int plspm_device_count(int32_t* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return fail(PLSPM_ERR_CUDA, cudaGetErrorString(e)); }
  *count = n;
  return PLSPM_OK;
}
int plspm_set_device(int32_t device) {
  CK(cudaSetDevice(device));
  return PLSPM_OK;
}

int plspm_model_create(int32_t L, const int32_t* block_sizes, const int8_t* modes, const int8_t* path, int32_t scaled,
                       int32_t tile_policy, plspm_model** out) {
  if (!block_sizes || !modes || !path || !out) return fail(PLSPM_ERR_INVALID, "null argument");
  plspm_model* m = new plspm_model();
  std::string err;
  if (build_model(L, block_sizes, modes, path, scaled, tile_policy, m->h, err)) {
    delete m;
    return fail(PLSPM_ERR_INVALID, err);
  }
  if ((size_t)m->h.Ppad * 8 > 64 * 1024) {
    delete m;
    return fail(PLSPM_ERR_UNSUPPORTED, "more than 8192 (padded) manifest variables");
  }
  CK(cudaGetDevice(&m->device));
  HostModel& h = m->h;
  ModelView v = h.host_view();
  int rc = 0;
  rc |= upload_vec(m, h.lv_off, &v.lv_off);
  rc |= upload_vec(m, h.lv_k, &v.lv_k);
  rc |= upload_vec(m, h.lv_mode, &v.lv_mode);
  rc |= upload_vec(m, h.col_lv, &v.col_lv);
  rc |= upload_vec(m, h.col_src, &v.col_src);
  rc |= upload_vec(m, h.path, &v.path);
  rc |= upload_vec(m, h.tile_sa, &v.tile_sa);
  rc |= upload_vec(m, h.tile_sb, &v.tile_sb);
  rc |= upload_vec(m, h.tile_of, &v.tile_of);
  rc |= upload_vec(m, h.lane_tile, &v.lane_tile);
  rc |= upload_vec(m, h.pair_l, &v.pair_l);
  rc |= upload_vec(m, h.pair_j, &v.pair_j);
  rc |= upload_vec(m, h.pair_voff, &v.pair_voff);
  rc |= upload_vec(m, h.lv_pair_begin, &v.lv_pair_begin);
  rc |= upload_vec(m, h.eff_from, &v.eff_from);
  rc |= upload_vec(m, h.eff_to, &v.eff_to);
  rc |= upload_vec(m, h.chol_b_off, &v.chol_b_off);
  rc |= upload_vec(m, h.pred_begin, &v.pred_begin);
  rc |= upload_vec(m, h.pred_idx, &v.pred_idx);
  rc |= upload_vec(m, h.succ_begin, &v.succ_begin);
  rc |= upload_vec(m, h.succ_idx, &v.succ_idx);
  rc |= upload_vec(m, h.omega, &v.omega);
  if (rc) {
    plspm_model_destroy(m);
    return PLSPM_ERR_CUDA;
  }
  m->dv = v;
  *out = m;
  return PLSPM_OK;
}

void plspm_model_destroy(plspm_model* m) {
  if (!m) return;
  for (void* p : m->dev_allocs) cudaFree(p);
  delete m;
}

int plspm_model_query(const plspm_model* m, int32_t* info) {
  if (!m || !info) return fail(PLSPM_ERR_INVALID, "null argument");
  std::memset(info, 0, 16 * sizeof(int32_t));
  const HostModel& h = m->h;
  info[0] = h.L; info[1] = h.P; info[2] = h.Ppad; info[3] = h.n_tiles; info[4] = h.n_tg; info[5] = h.n_pairs;
  info[6] = h.n_eff; info[7] = h.n_out(); info[8] = h.full; info[9] = h.scaled; info[10] = h.n_cross; info[11] = m->numeric ? 1 : 0;
  int nz = 0;  // pair-product columns of the tile set (rows / 6 of the int8 Gram GEMM)
  for (int t = 0; t < h.n_tiles; ++t)
    for (int r = 0; r < SLOT; ++r)
      for (int c = (h.tile_sa[t] == h.tile_sb[t] ? r : 0); c < SLOT; ++c)
        nz += (h.col_lv[h.tile_sa[t] * SLOT + r] >= 0 && h.col_lv[h.tile_sb[t] * SLOT + c] >= 0) ? 1 : 0;
  info[12] = nz;
  return PLSPM_OK;
}

int plspm_model_set_numeric(plspm_model* m, int32_t on) {
  if (!m) return fail(PLSPM_ERR_INVALID, "null argument");
  m->numeric = on != 0;
  return PLSPM_OK;
}

int plspm_model_effects(const plspm_model* m, int32_t* from, int32_t* to) {
  if (!m || !from || !to) return fail(PLSPM_ERR_INVALID, "null argument");
  for (int e = 0; e < m->h.n_eff; ++e) { from[e] = m->h.eff_from[e]; to[e] = m->h.eff_to[e]; }
  return PLSPM_OK;
}

// A data handle only depends on the column layout of the model it was created with (block sizes in
// path order); any model with the same layout -- other modes, paths, tile policy, treatment -- may use it.
static bool same_layout(const plspm_model* a, const plspm_model* b) {
  if (a == b) return true;
  return a && b && a->device == b->device && a->h.L == b->h.L && a->h.P == b->h.P && a->h.Ppad == b->h.Ppad &&
         a->h.lv_off == b->h.lv_off && a->h.lv_k == b->h.lv_k && a->h.col_src == b->h.col_src;
}

static int ws_reserve(plspm_data* d, size_t bytes) {
  if (d->ws.bytes >= bytes) return 0;
  if (d->ws.ptr) g_pool.release(d->ws.ptr);
  d->ws.ptr = nullptr;
  d->ws.bytes = 0;
  CK(g_pool.alloc(&d->ws.ptr, bytes));
  d->ws.bytes = bytes;
  return 0;
}

int plspm_data_create(const plspm_model* m, const double* X, int64_t N, int64_t ld, int32_t x_is_device,
                      plspm_data** out) {
  if (!m || !X || !out) return fail(PLSPM_ERR_INVALID, "null argument");
  const HostModel& h = m->h;
  if (N < 2) return fail(PLSPM_ERR_INVALID, "need at least two observations");
  if (N >= ((int64_t)1 << 32)) return fail(PLSPM_ERR_UNSUPPORTED, "more than 2^32-1 observations");
  if (ld < h.P) return fail(PLSPM_ERR_INVALID, "leading dimension smaller than the number of manifest variables");
  plspm_data* d = new plspm_data();
  d->model = m;
  d->N = N;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    delete d;
    return fail(PLSPM_ERR_CUDA, "no CUDA device: plspm_b200 has no CPU path");
  }
  d->sm_count = prop.multiProcessorCount;
  d->max_smem = (int)prop.sharedMemPerBlockOptin;
  auto bail = [&](int rc) { plspm_data_destroy(d); return rc; };
  static const bool tracing = getenv("PLSPM_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto trace = [&](const char* what) {
    if (!tracing) return;
    cudaStreamSynchronize(d->stream);
    fprintf(stderr, "[plspm_data_create] %-18s %8.3f ms\n", what,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  };
  if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(PLSPM_ERR_CUDA, "cudaStreamCreate failed"));
  cudaStream_t st = d->stream;
  trace("stream created");
  double* raw = nullptr;
  const double* Xd = X;
  int rc = 0;
  auto body = [&]() -> int {
    CK(g_pool.alloc((void**)&d->X, (size_t)N * h.Ppad * sizeof(double)));
    CK(g_pool.alloc((void**)&d->mu, (size_t)h.Ppad * sizeof(double)));
    CK(cudaMemsetAsync(d->mu, 0, (size_t)h.Ppad * sizeof(double), st));
    if (!x_is_device) {
      CK(g_pool.alloc((void**)&raw, (size_t)N * h.P * sizeof(double)));
      if (ld == h.P)
        CK(cudaMemcpyAsync(raw, X, (size_t)N * h.P * sizeof(double), cudaMemcpyHostToDevice, st));
      else
        CK(cudaMemcpy2DAsync(raw, (size_t)h.P * sizeof(double), X, (size_t)ld * sizeof(double),
                             (size_t)h.P * sizeof(double), (size_t)N, cudaMemcpyHostToDevice, st));
      trace("h2d issued");
      Xd = raw;
      ld = h.P;
    }
    const int nblocks = (int)std::min<int64_t>(1024, (N + 63) / 64);
    const int64_t rpb = (N + nblocks - 1) / nblocks;
    double* partial = nullptr;
    CK(g_pool.alloc((void**)&partial, (size_t)nblocks * h.P * sizeof(double)));
    int* src_col = nullptr;
    CK(g_pool.alloc((void**)&src_col, (size_t)h.P * sizeof(int)));
    CK(cudaMemcpyAsync(src_col, h.src_col.data(), (size_t)h.P * sizeof(int), cudaMemcpyHostToDevice, st));
    d->timer.begin(ST_UPLOAD, st);
    colsum_partial_kernel<<<nblocks, 256, 0, st>>>(Xd, N, ld, h.P, rpb, partial);
    d->timer.end(st);
    d->timer.begin(ST_UPLOAD, st);
    colmean_final_kernel<<<(h.P + 127) / 128, 128, 0, st>>>(partial, nblocks, h.P, N, src_col, d->mu);
    d->timer.end(st);
    d->timer.begin(ST_UPLOAD, st);
    relayout_kernel<<<d->sm_count * 8, 256, 0, st>>>(Xd, N, ld, h.Ppad, m->dv.col_src, d->mu, d->X);
    d->timer.end(st);
    CK(cudaGetLastError());
    trace("relayout done");
    // fp16 copy for the tensor-core sign vote of sparse tile sets (PLSPM_VOTE=exact disables it)
    static const bool vote_exact = getenv("PLSPM_VOTE") && std::string(getenv("PLSPM_VOTE")) == "exact";
    int nsl_pad_chk = 1;
    while (nsl_pad_chk * SLOT < h.kmax) nsl_pad_chk <<= 1;
    if (!h.full && !vote_exact && N >= 4096 && nsl_pad_chk <= 32 && h.L * nsl_pad_chk <= SG_THREADS) {
      double* sq = nullptr;
      CK(g_pool.alloc((void**)&sq, (size_t)nblocks * h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->inv_sd, (size_t)h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->Xh, (size_t)N * h.Ppad * sizeof(__half)));
      d->timer.begin(ST_UPLOAD, st);
      colsq_partial_kernel<<<nblocks, 256, 0, st>>>(d->X, N, h.Ppad, rpb, sq);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      inv_sd_kernel<<<(h.Ppad + 127) / 128, 128, 0, st>>>(sq, nblocks, h.Ppad, N, d->inv_sd);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      make_half_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N, h.Ppad, d->inv_sd, d->Xh);
      d->timer.end(st);
      CK(g_pool.alloc((void**)&d->Xf, (size_t)N * h.Ppad * sizeof(float)));
      d->timer.begin(ST_UPLOAD, st);
      make_float_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N * h.Ppad, d->Xf);
      d->timer.end(st);
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(st));
      g_pool.release(sq);
      d->blas = blas_acquire(dev);  // cublasCreate costs milliseconds: handles are recycled per device
      if (!d->blas) return fail(PLSPM_ERR_CUDA, "cublasCreate failed");
      d->blas_device = dev;
      cublasSetStream(d->blas, st);
      d->fast_vote = true;
      trace("fp16 copy + blas");
    }
    if (!d->Xf && N >= 4096) {  // fp32 copy for the score screening of the non-metric criterion pass
      CK(g_pool.alloc((void**)&d->Xf, (size_t)N * h.Ppad * sizeof(float)));
      d->timer.begin(ST_UPLOAD, st);
      make_float_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N * h.Ppad, d->Xf);
      d->timer.end(st);
      CK(cudaGetLastError());
    }
    // integer digit planes for the tensor-core column sums (PLSPM_COLSUM=fp64 keeps the fp64 kernel)
    static const bool colsum_fp64 = getenv("PLSPM_COLSUM") && std::string(getenv("PLSPM_COLSUM")) == "fp64";
    if (!colsum_fp64 && N >= 4096 && N < ((int64_t)1 << 31) - 16) {
      if (!d->blas) {
        d->blas = blas_acquire(dev);
        if (!d->blas) return fail(PLSPM_ERR_CUDA, "cublasCreate failed");
        d->blas_device = dev;
        cublasSetStream(d->blas, st);
      }
      d->Npad = (N + 15) / 16 * 16;
      double *amax = nullptr, *qscale = nullptr;
      CK(g_pool.alloc((void**)&amax, (size_t)nblocks * h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&qscale, (size_t)h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->dscale, (size_t)h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->D8, (size_t)I8_DIGITS * h.Ppad * d->Npad));
      d->timer.begin(ST_UPLOAD, st);
      colabsmax_partial_kernel<<<nblocks, 256, 0, st>>>(d->X, N, h.Ppad, rpb, amax);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      digit_scale_kernel<<<(h.Ppad + 127) / 128, 128, 0, st>>>(amax, nblocks, h.Ppad, d->dscale, qscale);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      digits_kernel<<<dim3((h.Ppad + 31) / 32, (unsigned)((d->Npad + 127) / 128)), 256, 0, st>>>(d->X, N, h.Ppad, d->Npad,
                                                                                            qscale, d->D8);
      d->timer.end(st);
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(st));
      g_pool.release(amax);
      d->i8_colsum = true;
      trace("digit planes");
      // pair-product planes of the model's tile set, if they fit the budget (PLSPM_I8_GRAM_GB, default 24)
      static const double z_budget_gb = getenv("PLSPM_I8_GRAM_GB") ? atof(getenv("PLSPM_I8_GRAM_GB")) : 24.0;
      std::vector<int> zp, zq, zd1, zd2;
      for (int t = 0; t < h.n_tiles; ++t) {
        const int sa = h.tile_sa[t], sb = h.tile_sb[t];
        for (int r = 0; r < SLOT; ++r)
          for (int c = (sa == sb ? r : 0); c < SLOT; ++c) {
            const int pp = sa * SLOT + r, qq = sb * SLOT + c;
            if (h.col_lv[pp] < 0 || h.col_lv[qq] < 0) continue;  // padding columns are zero: their moments stay 0
            zp.push_back(pp); zq.push_back(qq);
            zd1.push_back(t * TILE + r * SLOT + c);
            zd2.push_back(sa == sb && r != c ? t * TILE + c * SLOT + r : -1);
          }
      }
      const double z_bytes = (double)I8_DIGITS * zp.size() * d->Npad;
      static const double z_chunk_gb = getenv("PLSPM_I8_CHUNK_GB") ? atof(getenv("PLSPM_I8_CHUNK_GB")) : 12.0;
      if (!zp.empty() && z_budget_gb > 0 && (int64_t)I8_DIGITS * (int64_t)zp.size() < ((int64_t)1 << 31)) {
        const int nz = (int)zp.size();
        CK(g_pool.alloc((void**)&d->zp, (size_t)nz * 4));
        CK(g_pool.alloc((void**)&d->zq, (size_t)nz * 4));
        CK(g_pool.alloc((void**)&d->zdst, (size_t)nz * 4));
        CK(g_pool.alloc((void**)&d->zdst2, (size_t)nz * 4));
        CK(g_pool.alloc((void**)&d->zqscale, (size_t)nz * 8));
        CK(g_pool.alloc((void**)&d->zdscale, (size_t)nz * 8));
        CK(cudaMemcpyAsync(d->zp, zp.data(), (size_t)nz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->zq, zq.data(), (size_t)nz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->zdst, zd1.data(), (size_t)nz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->zdst2, zd2.data(), (size_t)nz * 4, cudaMemcpyHostToDevice, st));
        d->timer.begin(ST_UPLOAD, st);
        zscale_kernel<<<(nz + 127) / 128, 128, 0, st>>>(nz, d->zp, d->zq, qscale, d->dscale, d->zqscale, d->zdscale);
        d->timer.end(st);
        CK(cudaGetLastError());
        if (z_bytes <= z_budget_gb * 1e9) {  // resident planes: generated once
          CK(g_pool.alloc((void**)&d->Z8, (size_t)I8_DIGITS * nz * d->Npad));
          d->timer.begin(ST_UPLOAD, st);
          zdigits_kernel<<<dim3((nz + 31) / 32, (unsigned)((d->Npad + 127) / 128)), 256, 0, st>>>(
              d->X, N, h.Ppad, 0, d->Npad, nz, d->zp, d->zq, d->zqscale, d->Z8);
          d->timer.end(st);
          CK(cudaGetLastError());
        } else {  // too large to keep: a chunk of rows is generated per batch, right before its GEMM
          int64_t rows = (int64_t)(z_chunk_gb * 1e9 / ((double)I8_DIGITS * nz));
          rows = std::min<int64_t>(rows / 128 * 128, I8_KCHUNK);
          d->z_chunk_rows = std::max<int64_t>(rows, 4096);
        }
        CK(cudaStreamSynchronize(st));  // (the host vectors above are pageable: copies are done by now)
        d->n_zcols = nz;
        d->z_model = m;
        trace("pair-product planes");
      }
      g_pool.release(qscale);
    }
    CK(cudaStreamSynchronize(st));
    d->timer.collect();
    trace("timer collected");
    {  // a NaN / Inf anywhere in a column poisons its mean: one P-length check instead of a host pass over X
      std::vector<double> mu_host(h.Ppad);
      CK(cudaMemcpy(mu_host.data(), d->mu, (size_t)h.Ppad * sizeof(double), cudaMemcpyDeviceToHost));
      for (double v : mu_host)
        if (!std::isfinite(v)) return fail(PLSPM_ERR_INVALID, "non-finite values in the observation matrix");
    }
    g_pool.release(partial);
    g_pool.release(src_col);
    return 0;
  };
  rc = body();
  if (raw) g_pool.release(raw);
  if (rc) return bail(rc);
  *out = d;
  return PLSPM_OK;
}

void plspm_data_destroy(plspm_data* d) {
  if (!d) return;
  if (d->X) g_pool.release(d->X);
  if (d->mu) g_pool.release(d->mu);
  if (d->Xh) g_pool.release(d->Xh);
  if (d->Xf) g_pool.release(d->Xf);
  if (d->inv_sd) g_pool.release(d->inv_sd);
  if (d->D8) g_pool.release(d->D8);
  if (d->Z8) g_pool.release(d->Z8);
  if (d->zdscale) g_pool.release(d->zdscale);
  if (d->zp) g_pool.release(d->zp);
  if (d->zq) g_pool.release(d->zq);
  if (d->zqscale) g_pool.release(d->zqscale);
  if (d->zdst) g_pool.release(d->zdst);
  if (d->zdst2) g_pool.release(d->zdst2);
  if (d->dscale) g_pool.release(d->dscale);
  if (d->blas) blas_release(d->blas_device, d->blas);
  if (d->ws.ptr) g_pool.release(d->ws.ptr);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

// ------------------------------------------------------------------------------------------------
// launch plans, workspace layout, batch driver
// ------------------------------------------------------------------------------------------------
struct StreamPlan {  // one streaming pass over X (Gram tiles or cross-moment tiles)
  int RT = 0, stages = 0, n_chunks = 1;
  int64_t chunk_rows = 0, n_groups = 0;
  size_t smem = 0;
};
struct BatchPlan {
  StreamPlan gram, cross;
  int cs_chunks = 1;  // row chunks of the column-sum kernel
  int64_t cs_chunk_rows = 0;
  // criterion pass of the numeric non-metric path
  int cv_nsl_pad = 1, cv_reps_per_cta = 1, cv_rows = 1;
  bool cv_f32 = false;
  unsigned cv_gx = 1, cv_gy = 1;
  size_t cv_smem = 0;
};

static void plan_colsum(const plspm_data* d, int64_t nb, BatchPlan& g) {
  const HostModel& h = d->model->h;
  const int64_t blocks_xy = (int64_t)((h.Ppad + CS_COLS - 1) / CS_COLS) * ((nb + CS_REPS - 1) / CS_REPS);
  const int64_t want = (int64_t)d->sm_count * 8;
  int64_t chunks = std::max<int64_t>(1, (want + blocks_xy - 1) / blocks_xy);
  chunks = std::min<int64_t>(chunks, std::max<int64_t>(1, d->N / (CS_ROWS * 8)));
  chunks = std::min<int64_t>(chunks, 65535);
  int64_t rows = ((d->N + chunks - 1) / chunks + CS_ROWS - 1) / CS_ROWS * CS_ROWS;
  g.cs_chunk_rows = rows;
  g.cs_chunks = (int)((d->N + rows - 1) / rows);
}

// Ring geometry + row chunking of a streaming pass.  extra_row_bytes / extra_fixed_bytes: per-CTA
// shared memory the kernel needs per ring row and in total besides the ring (cross mode scratch).
static int plan_stream(const plspm_data* d, int64_t n_items, size_t extra_row_bytes, size_t extra_fixed_bytes,
                       StreamPlan& g) {
  const HostModel& h = d->model->h;
  const size_t row_bytes = (size_t)h.Ppad * 8;
  // stages of <= 64 KB / 32 rows; measured on c3: 32-row stages beat 16-row stages by 10 % (per-tile
  // handshakes amortise), three stages are enough once the producer refills opportunistically
  static const int stage_kb = getenv("PLSPM_GRAM_STAGE_KB") ? atoi(getenv("PLSPM_GRAM_STAGE_KB")) : 64;
  static const int want_stages = getenv("PLSPM_GRAM_STAGES") ? atoi(getenv("PLSPM_GRAM_STAGES")) : 3;
  const size_t budget = (size_t)d->max_smem - 8 * 1024;  // static shared memory (barriers, row lists) + slack
  if (budget < extra_fixed_bytes + 2 * (row_bytes + extra_row_bytes))
    return fail(PLSPM_ERR_UNSUPPORTED, "manifest rows too wide for the shared-memory ring");
  // Rows per stage matter more than the number of stages (every tile costs a list build, a pipeline
  // prologue/epilogue and, for odd counts, one padded row): measured on P=1024, 2 stages x 12 rows beat
  // 3 x 8 by 14 % and 6 x 4 by 41 %.  So: three stages if they still hold >= 16 rows each, else two.
  static const bool stage_env = getenv("PLSPM_GRAM_STAGE_KB") != nullptr;
  auto rows_for = [&](int st) {
    int64_t r = (int64_t)((budget - extra_fixed_bytes) / ((size_t)st * row_bytes + extra_row_bytes));
    if (stage_env || st >= 3) r = std::min<int64_t>(r, std::max<int64_t>(1, (int64_t)((size_t)stage_kb * 1024 / row_bytes)));
    return std::min<int64_t>(r, 32);
  };
  int stages = std::max(2, std::min(want_stages, GRAM_MAX_STAGES));
  int64_t RT = rows_for(stages);
  if (stages > 2 && RT < 16 && rows_for(2) > RT) {
    stages = 2;
    RT = rows_for(2);
  }
  if (RT < 1) return fail(PLSPM_ERR_UNSUPPORTED, "manifest rows too wide for the shared-memory ring");
  const int64_t n_tiles_rt = (d->N + RT - 1) / RT;
  stages = (int)std::min<int64_t>(stages, std::max<int64_t>(2, n_tiles_rt));
  const int64_t n_groups = (n_items + GRAM_WARPS - 1) / GRAM_WARPS;
  // split rows only when there are too few (replicate, tile group) items to fill the chip
  int64_t n_chunks = 1;
  const int64_t want = (int64_t)d->sm_count * 2;
  if (n_groups < want) n_chunks = (want + n_groups - 1) / n_groups;
  n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_chunks, n_tiles_rt / 4 > 0 ? n_tiles_rt / 4 : 1));
  int64_t chunk_rows = ((n_tiles_rt + n_chunks - 1) / n_chunks) * RT;
  n_chunks = (d->N + chunk_rows - 1) / chunk_rows;
  g.RT = (int)RT; g.stages = stages; g.n_chunks = (int)n_chunks; g.chunk_rows = chunk_rows; g.n_groups = n_groups;
  g.smem = (size_t)stages * RT * row_bytes + extra_fixed_bytes + (size_t)RT * extra_row_bytes;
  return 0;
}

static int plan_batch(const plspm_data* d, int64_t nb, BatchPlan& bp) {
  const HostModel& h = d->model->h;
  if (int rc = plan_stream(d, nb * h.n_tg, 0, 0, bp.gram)) return rc;
  if (!h.full && !d->model->numeric) {
    const size_t per_row = (size_t)GRAM_WARPS * h.ng * SLOT * 8;  // score scratch row of every warp
    if (int rc = plan_stream(d, nb * h.n_tg_cross, per_row, 8 * per_row, bp.cross)) return rc;
  }
  plan_colsum(d, nb, bp);
  if (d->model->numeric) {
    int nsl_pad = 1;
    while (nsl_pad * SLOT < h.kmax) nsl_pad <<= 1;
    if (nsl_pad > 32 || h.L * nsl_pad > SG_THREADS)
      return fail(PLSPM_ERR_UNSUPPORTED, "numeric non-metric path: L x (padded slots per block) exceeds 256");
    const int nbl = SG_THREADS / (h.L * nsl_pad);
    bp.cv_nsl_pad = nsl_pad;
    // the fp32-screened variant is correct but not yet faster than the fp64 one (12.6 vs ~10 ms per pass on c3:
    // branchy inner loop, 128-register cap); opt-in until it is tuned
    static const bool conv_f32 = getenv("PLSPM_CONV_F32") && atoi(getenv("PLSPM_CONV_F32")) != 0;
    bp.cv_f32 = conv_f32 && d->Xf != nullptr;
    const size_t esz = bp.cv_f32 ? 4 : 8;
    const int rpt = bp.cv_f32 ? CvTraits<float>::RPT : CvTraits<double>::RPT;
    bp.cv_reps_per_cta = nbl * rpt;
    const size_t fixed = (size_t)SG_THREADS * rpt * 8 + 16;
    // at most 64 rows per staged tile, and small enough for two CTAs per SM
    bp.cv_rows = (int)std::max<size_t>(1, std::min<size_t>(bp.cv_f32 ? 64 : 32, ((size_t)d->max_smem / 2 - 8192 - fixed) /
                                                                                 (((size_t)h.Ppad + bp.cv_reps_per_cta) * esz)));
    bp.cv_smem = (size_t)bp.cv_rows * ((size_t)h.Ppad + bp.cv_reps_per_cta) * esz + fixed;
    bp.cv_gy = (unsigned)((nb + bp.cv_reps_per_cta - 1) / bp.cv_reps_per_cta);
    bp.cv_gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((d->N + bp.cv_rows - 1) / bp.cv_rows,
                                                                 (4 * d->sm_count + bp.cv_gy - 1) / bp.cv_gy));
  }
  return 0;
}

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Device workspace of one batch of up to nb replicates (offsets into plspm_data::ws).
constexpr int FAST_RC = 4096;  // rows per tensor-core GEMM chunk (bounds the fp32 accumulation error)
struct BatchBuffers {
  size_t total = 0;
  size_t counts, idx, G, Gpart, colsum, cspart, ws, out, iters, status, wf, CG, CGpart, sh, BT, Cf, rep_map;
  // single-fit outputs
  size_t weights, loadings, r2, paths, totalfx, crossl, coef, shift, scores;
  // numeric non-metric path: per-replicate iteration state
  size_t num_a, num_co, num_cn, num_so, num_sn, num_meta, num_done, num_cpart, num_cmain;
  // tensor-core column sums: int8 multiplicities, int32 digit sums, overflow flag
  size_t c8, s32, ovf, zs32, zchunk;
};
static BatchBuffers layout_batch(const plspm_data* d, int64_t nb, const BatchPlan& bp, bool with_counts, bool with_idx,
                                 bool rows_on_device_of_caller, bool single_fit, bool want_scores) {
  const HostModel& h = d->model->h;
  BatchBuffers b;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(std::max<size_t>(bytes, 8)); return o; };
  const size_t gsz = (size_t)h.n_tiles * TILE * 8, csz = (size_t)h.n_cross * TILE * 8;
  b.counts = take(with_counts ? (size_t)nb * d->N * 4 : 0);
  b.idx = take(with_idx ? (size_t)nb * d->N * 4 : 0);
  b.G = take((size_t)nb * gsz);
  b.Gpart = take(bp.gram.n_chunks > 1 ? (size_t)nb * gsz * bp.gram.n_chunks : 0);
  b.colsum = take((size_t)nb * h.Ppad * 8);
  b.cspart = take(bp.cs_chunks > 1 ? (size_t)nb * h.Ppad * 8 * bp.cs_chunks : 0);
  b.ws = take((size_t)nb * h.ws_doubles * 8);
  b.out = take(rows_on_device_of_caller || single_fit ? 0 : (size_t)nb * h.n_out() * 8);
  b.iters = take((size_t)nb * 4);
  b.status = take((size_t)nb * 4);
  const bool numeric = d->model->numeric;
  const bool vote = !h.full && !numeric;  // sparse tile set: buffers of the cross-moment / sign-vote pass
  b.wf = take(vote ? (size_t)nb * h.Ppad * 8 : 0);
  b.CG = take(vote ? (size_t)nb * csz : 0);
  b.CGpart = take(vote && bp.cross.n_chunks > 1 ? (size_t)nb * csz * bp.cross.n_chunks : 0);
  b.sh = take(vote ? (size_t)nb * h.L * 8 : 0);
  const bool fast = vote && d->fast_vote && !single_fit;
  b.num_a = take(numeric ? (size_t)nb * h.Ppad * 8 : 0);
  b.num_co = take(numeric ? (size_t)nb * h.Ppad * 8 : 0);
  b.num_cn = take(numeric ? (size_t)nb * h.Ppad * 8 : 0);
  b.num_so = take(numeric ? (size_t)nb * h.L * 8 : 0);
  b.num_sn = take(numeric ? (size_t)nb * h.L * 8 : 0);
  b.num_meta = take(numeric ? (size_t)nb * 16 : 0);
  b.num_done = take(8);
  const bool i8 = d->i8_colsum && with_counts;
  b.c8 = take(i8 ? (size_t)nb * d->Npad : 0);
  b.s32 = take(i8 ? (size_t)nb * I8_DIGITS * h.Ppad * sizeof(int32_t) : 0);
  b.ovf = take(8);
  b.zs32 = take(i8 && d->n_zcols ? (size_t)nb * I8_DIGITS * d->n_zcols * sizeof(int32_t) : 0);
  b.zchunk = take(i8 && d->n_zcols && !d->Z8 ? (size_t)I8_DIGITS * d->n_zcols * d->z_chunk_rows : 0);
  b.num_cpart = take(numeric ? (size_t)nb * bp.cv_gx * 8 : 0);
  b.num_cmain = take(numeric ? (size_t)nb * 8 : 0);
  const size_t ldl = (size_t)(nb + 7) / 8 * 8;  // replicate stride of the LV-major score / cross-moment layout
  b.BT = take(fast ? ldl * h.L * FAST_RC * sizeof(__half) : 0);
  b.Cf = take(fast ? ldl * h.L * h.Ppad * sizeof(float) : 0);
  b.rep_map = take((size_t)nb * 4);
  if (single_fit) {
    const size_t L = h.L, P = h.P;
    b.weights = take(P * 8); b.loadings = take(P * 8); b.r2 = take(L * 8); b.paths = take(L * L * 8);
    b.totalfx = take(L * L * 8); b.crossl = take(P * L * 8); b.coef = take((size_t)h.Ppad * 8); b.shift = take(L * 8);
    b.scores = take(want_scores ? (size_t)d->N * L * 8 : 0);
  }
  b.total = off;
  return b;
}

static int launch_stream(plspm_data* d, bool cross, int64_t nb, const uint32_t* counts_dev, const StreamPlan& sp,
                         double* out_tiles, double* part_tiles, const double* wf, const int* rep_map = nullptr) {
  const plspm_model* m = d->model;
  const HostModel& h = m->h;
  cudaStream_t st = d->stream;
  GramParams p;
  p.X = d->X; p.counts = counts_dev; p.N = d->N; p.Ppad = h.Ppad;
  p.n_tiles = cross ? h.n_cross : h.n_tiles;
  p.n_tg = cross ? h.n_tg_cross : h.n_tg;
  p.tile_sa = m->dv.tile_sa; p.tile_sb = m->dv.tile_sb; p.lane_tile = m->dv.lane_tile;
  p.n_items = nb * p.n_tg; p.n_chunks = sp.n_chunks; p.chunk_rows = sp.chunk_rows; p.RT = sp.RT; p.stages = sp.stages;
  p.G = (sp.n_chunks > 1) ? part_tiles : out_tiles;
  p.L = h.L; p.ng = h.ng; p.lv_off = m->dv.lv_off; p.lv_k = m->dv.lv_k; p.wf = wf; p.rep_map = rep_map;
  const int64_t n_groups = (p.n_items + GRAM_WARPS - 1) / GRAM_WARPS;
  const int64_t grid = n_groups * sp.n_chunks;
  if (grid > 0x7fffffff) return fail(PLSPM_ERR_UNSUPPORTED, "batch too large for one launch");
  d->timer.begin(cross ? ST_CROSS : ST_GRAM, st);
  if (cross) {
    CK(cudaFuncSetAttribute(gram_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem));
    gram_kernel<true><<<(unsigned)grid, GRAM_THREADS, sp.smem, st>>>(p);
  } else {
    CK(cudaFuncSetAttribute(gram_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem));
    gram_kernel<false><<<(unsigned)grid, GRAM_THREADS, sp.smem, st>>>(p);
  }
  d->timer.end(st);
  CK(cudaGetLastError());
  if (sp.n_chunks > 1) {
    d->timer.begin(ST_REDUCE, st);
    reduce_chunks_kernel<<<d->sm_count * 4, 256, 0, st>>>(part_tiles, nb, sp.n_chunks, (int64_t)p.n_tiles * TILE,
                                                          out_tiles);
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  return 0;
}

// Exact redo (fp64 cross moments) of the replicates the low-precision vote could not decide.
static int redo_exact(plspm_data* d, int64_t n_list, const int* rep_map_dev, const uint32_t* counts_dev,
                      const BatchBuffers& bb, int scheme, double tol, int max_iter, const BatchPlan& bp,
                      double* out_rows) {
  const plspm_model* m = d->model;
  const HostModel& h = m->h;
  cudaStream_t st = d->stream;
  char* base = (char*)d->ws.ptr;
  auto D = [&](size_t o) { return (double*)(base + o); };
  StreamPlan sp = bp.cross;
  sp.n_chunks = 1;  // partial buffers are laid out per replicate position; keep whole-row passes here
  sp.chunk_rows = ((d->N + sp.RT - 1) / sp.RT) * sp.RT;
  if (int rc = launch_stream(d, true, n_list, counts_dev, sp, D(bb.CG), D(bb.CGpart), D(bb.wf), rep_map_dev)) return rc;
  SolveBatch b;
  std::memset(&b, 0, sizeof(b));
  b.M = m->dv;
  b.G = D(bb.G); b.g_stride = (int64_t)h.n_tiles * TILE;
  b.colsum = D(bb.colsum); b.cs_stride = h.Ppad;
  b.mu = d->mu; b.N = (double)d->N; b.scheme = scheme; b.tol = tol; b.max_iter = max_iter;
  b.ws = D(bb.ws);
  b.iters = (int*)(base + bb.iters); b.status = (int*)(base + bb.status);
  b.cross = D(bb.CG); b.cross_stride = (int64_t)h.n_cross * TILE;
  b.phase = 2; b.rep_map = rep_map_dev;
  b.out_rows = out_rows; b.out_stride = h.n_out();
  const size_t smem = h.solver_smem_doubles() * sizeof(double);
  d->timer.begin(ST_SOLVE, st);
  solve_kernel<<<(unsigned)n_list, SOLVE_THREADS, smem, st>>>(b);
  d->timer.end(st);
  CK(cudaGetLastError());
  return 0;
}

// First and second moments of every replicate of a batch: Gram tiles + column sums.
static int launch_moments(plspm_data* d, int64_t nb, const uint32_t* counts_dev, const BatchBuffers& bb,
                          const BatchPlan& bp) {
  const HostModel& h = d->model->h;
  cudaStream_t st = d->stream;
  char* base = (char*)d->ws.ptr;
  auto D = [&](size_t o) { return (double*)(base + o); };
  const bool i8 = d->i8_colsum && counts_dev;
  int8_t* c8 = (int8_t*)(base + bb.c8);
  if (i8) {
    d->timer.begin(ST_COLSUM, st);
    counts8_kernel<<<d->sm_count * 8, 256, 0, st>>>(counts_dev, d->N, d->Npad, nb, c8, (int*)(base + bb.ovf));
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  bool gram_done = false;
  if (i8 && d->n_zcols && d->z_model == d->model) {
    // all Gram tiles of the batch: counts8 [nb x N] x pair-product planes [N x 6 n_zcols], exact int32 sums
    int32_t* zs = (int32_t*)(base + bb.zs32);
    const int32_t one = 1, zero = 0;
    const int gemm_m = I8_DIGITS * d->n_zcols;
    const int64_t g_stride = (int64_t)h.n_tiles * TILE;
    const int64_t step = d->Z8 ? I8_KCHUNK : d->z_chunk_rows;
    CK(cudaMemsetAsync(D(bb.G), 0, (size_t)nb * g_stride * 8, st));
    gram_done = true;
    for (int64_t k0 = 0; k0 < d->Npad; k0 += step) {
      const int kc = (int)std::min<int64_t>(step, d->Npad - k0);
      const int8_t* planes = d->Z8 ? d->Z8 + k0 : (const int8_t*)(base + bb.zchunk);
      const int64_t ld = d->Z8 ? d->Npad : kc;
      if (!d->Z8) {
        d->timer.begin(ST_GRAM_I8, st);
        zdigits_kernel<<<dim3((d->n_zcols + 31) / 32, (unsigned)((kc + 127) / 128)), 256, 0, st>>>(
            d->X, d->N, h.Ppad, k0, kc, d->n_zcols, d->zp, d->zq, d->zqscale, (int8_t*)(base + bb.zchunk));
        d->timer.end(st);
        CK(cudaGetLastError());
      }
      d->timer.begin(ST_GRAM_I8, st);
      cublasStatus_t cs = cublasGemmEx(d->blas, CUBLAS_OP_T, CUBLAS_OP_N, gemm_m, (int)nb, kc, &one, planes, CUDA_R_8I,
                                       (int)ld, c8 + k0, CUDA_R_8I, (int)d->Npad, &zero, zs, CUDA_R_32I, gemm_m,
                                       CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
      d->timer.end(st);
      if (cs != CUBLAS_STATUS_SUCCESS) { gram_done = false; break; }
      d->timer.begin(ST_GRAM_I8, st);
      zcombine_kernel<<<d->sm_count * 8, 256, 0, st>>>(zs, nb, d->n_zcols, d->zdscale, d->zdst, d->zdst2, k0 > 0 ? 1 : 0,
                                                      g_stride, D(bb.G));
      d->timer.end(st);
      CK(cudaGetLastError());
    }
    if (!gram_done) d->z_model = nullptr;  // no int8 GEMM for this shape: fp64 kernel from now on
  }
  if (!gram_done)
    if (int rc = launch_stream(d, false, nb, counts_dev, bp.gram, D(bb.G), D(bb.Gpart), nullptr)) return rc;
  if (i8) {
    int32_t* s32 = (int32_t*)(base + bb.s32);
    const int32_t one = 1, zero = 0;
    const int gemm_m = I8_DIGITS * h.Ppad;
    bool ok = true;
    for (int64_t k0 = 0; k0 < d->Npad && ok; k0 += I8_KCHUNK) {
      const int kc = (int)std::min<int64_t>(I8_KCHUNK, d->Npad - k0);
      d->timer.begin(ST_COLSUM, st);
      cublasStatus_t cs = cublasGemmEx(d->blas, CUBLAS_OP_T, CUBLAS_OP_N, gemm_m, (int)nb, kc, &one, d->D8 + k0, CUDA_R_8I,
                                       (int)d->Npad, c8 + k0, CUDA_R_8I, (int)d->Npad, &zero, s32, CUDA_R_32I, gemm_m,
                                       CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
      d->timer.end(st);
      if (cs != CUBLAS_STATUS_SUCCESS) { ok = false; break; }
      d->timer.begin(ST_COLSUM, st);
      digits_combine_kernel<<<d->sm_count * 2, 256, 0, st>>>(s32, nb, h.Ppad, d->dscale, k0 > 0 ? 1 : 0, D(bb.colsum));
      d->timer.end(st);
      CK(cudaGetLastError());
    }
    if (ok) return 0;
    d->i8_colsum = false;  // this cuBLAS build has no int8 GEMM for the shape: fp64 kernel from now on
  }
  dim3 grid_cs((h.Ppad + CS_COLS - 1) / CS_COLS, (unsigned)((nb + CS_REPS - 1) / CS_REPS), bp.cs_chunks);
  d->timer.begin(ST_COLSUM, st);
  const size_t cs_smem_bytes = (size_t)(CS_ROWS * CS_COLS + CS_ROWS * CS_REPS) * 8;
  CK(cudaFuncSetAttribute(colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)cs_smem_bytes));
  colsum_kernel<<<grid_cs, 256, cs_smem_bytes, st>>>(d->X, counts_dev, d->N, h.Ppad, nb, bp.cs_chunks, bp.cs_chunk_rows,
                                                     bp.cs_chunks > 1 ? D(bb.cspart) : D(bb.colsum));
  d->timer.end(st);
  if (bp.cs_chunks > 1) {
    d->timer.begin(ST_REDUCE, st);
    reduce_chunks_kernel<<<d->sm_count, 256, 0, st>>>(D(bb.cspart), nb, bp.cs_chunks, h.Ppad, D(bb.colsum));
    d->timer.end(st);
  }
  CK(cudaGetLastError());
  return 0;
}

// Numeric non-metric path (solver_num.h): the stopping rule is evaluated on the scores, so the outer
// iteration is driven from here -- one num_step launch (all unfinished replicates advance by one
// iteration, or finish) and one criterion pass over X per iteration.
static int run_batch_num(plspm_data* d, int64_t nb, const uint32_t* counts_dev, const BatchBuffers& bb, int scheme,
                         double tol, int max_iter, const BatchPlan& bp, double* out_rows, bool single_fit) {
  const plspm_model* m = d->model;
  const HostModel& h = m->h;
  cudaStream_t st = d->stream;
  char* base = (char*)d->ws.ptr;
  auto D = [&](size_t o) { return (double*)(base + o); };
  if (int rc = launch_moments(d, nb, counts_dev, bb, bp)) return rc;
  NumBatch b;
  std::memset(&b, 0, sizeof(b));
  b.M = m->dv;
  b.G = D(bb.G); b.g_stride = (int64_t)h.n_tiles * TILE;
  b.colsum = D(bb.colsum);
  b.N = (double)d->N; b.scheme = scheme; b.tol = tol; b.max_iter = max_iter;
  b.conv_part = D(bb.num_cpart); b.n_conv_part = (int)bp.cv_gx;
  b.conv_main = D(bb.num_cmain);
  b.ws = D(bb.ws);
  b.a = D(bb.num_a); b.coef_old = D(bb.num_co); b.coef_new = D(bb.num_cn);
  b.shift_old = D(bb.num_so); b.shift_new = D(bb.num_sn);
  b.meta = (int*)(base + bb.num_meta); b.n_done = (int*)(base + bb.num_done);
  b.out_rows = out_rows; b.out_stride = h.n_out();
  b.iters = (int*)(base + bb.iters); b.status = (int*)(base + bb.status);
  if (single_fit) {
    b.weights = D(bb.weights); b.loadings = D(bb.loadings); b.r2 = D(bb.r2); b.paths = D(bb.paths);
    b.total = D(bb.totalfx); b.crossloadings = D(bb.crossl); b.score_coef = D(bb.coef); b.score_shift = D(bb.shift);
  }
  CK(cudaMemsetAsync(b.meta, 0, (size_t)nb * 16, st));
  CK(cudaMemsetAsync(b.n_done, 0, 8, st));
  CK(cudaMemsetAsync(D(bb.num_cpart), 0, (size_t)nb * bp.cv_gx * 8, st));
  CK(cudaMemsetAsync(D(bb.num_cmain), 0, (size_t)nb * 8, st));
  const size_t smem = h.solver_smem_doubles() * sizeof(double);
  if (smem > (size_t)d->max_smem) return fail(PLSPM_ERR_UNSUPPORTED, "model too large for the solver's shared memory");
  CK(cudaFuncSetAttribute(num_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  CK(cudaFuncSetAttribute(conv_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.cv_smem));
  CK(cudaFuncSetAttribute(conv_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bp.cv_smem));
  const unsigned gy = (unsigned)((nb + bp.cv_reps_per_cta - 1) / bp.cv_reps_per_cta);
  int done = 0;
  for (int step = 0; step < max_iter + 4; ++step) {
    d->timer.begin(ST_SOLVE, st);
    num_step_kernel<<<(unsigned)nb, 128, smem, st>>>(b);
    d->timer.end(st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&done, b.n_done, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (done >= nb) break;
    d->timer.begin(ST_CONV, st);
    if (bp.cv_f32)
      conv_kernel<float><<<dim3(bp.cv_gx, gy), SG_THREADS, bp.cv_smem, st>>>(
          d->Xf, d->X, counts_dev, b.coef_old, b.coef_new, b.shift_old, b.shift_new, b.meta, d->N, h.Ppad, h.L,
          m->dv.lv_off, m->dv.lv_k, bp.cv_nsl_pad, bp.cv_rows, nb, D(bb.num_cpart));
    else
      conv_kernel<double><<<dim3(bp.cv_gx, gy), SG_THREADS, bp.cv_smem, st>>>(
          d->X, d->X, counts_dev, b.coef_old, b.coef_new, b.shift_old, b.shift_new, b.meta, d->N, h.Ppad, h.L,
          m->dv.lv_off, m->dv.lv_k, bp.cv_nsl_pad, bp.cv_rows, nb, D(bb.num_cpart));
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  if (done < nb) return fail(PLSPM_ERR_CUDA, "numeric non-metric iteration did not terminate");
  return 0;
}

// Shared implementation of fit (counts == null, one "replicate") and bootstrap batches.
static int run_batch(plspm_data* d, int64_t nb, const uint32_t* counts_dev, const BatchBuffers& bb, int scheme,
                     double tol, int max_iter, const BatchPlan& bp, double* out_rows, bool single_fit) {
  const plspm_model* m = d->model;
  const HostModel& h = m->h;
  const bool use_fast = !h.full && d->fast_vote && !single_fit && (int64_t)(nb + 8) * h.L < (1 << 30);
  cudaStream_t st = d->stream;
  char* base = (char*)d->ws.ptr;
  auto D = [&](size_t o) { return (double*)(base + o); };
  if (int rc = launch_moments(d, nb, counts_dev, bb, bp)) return rc;
  SolveBatch b;
  std::memset(&b, 0, sizeof(b));
  b.M = m->dv;
  b.G = D(bb.G); b.g_stride = (int64_t)h.n_tiles * TILE;
  b.colsum = D(bb.colsum); b.cs_stride = h.Ppad;
  b.mu = d->mu; b.N = (double)d->N; b.scheme = scheme; b.tol = tol; b.max_iter = max_iter;
  b.ws = D(bb.ws);
  b.iters = (int*)(base + bb.iters); b.status = (int*)(base + bb.status);
  b.wf = h.full ? nullptr : D(bb.wf);
  b.cross = h.full ? nullptr : D(bb.CG); b.cross_stride = (int64_t)h.n_cross * TILE;
  const size_t smem = h.solver_smem_doubles() * sizeof(double);
  if (smem > (size_t)d->max_smem) return fail(PLSPM_ERR_UNSUPPORTED, "model too large for the solver's shared memory");
  CK(cudaFuncSetAttribute(solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (!h.full) {
    // sparse tile set: final weights first, then the P x L cross-moment pass for the sign vote
    b.phase = 1;
    b.sh = D(bb.sh);
    d->timer.begin(ST_SOLVE, st);
    solve_kernel<<<(unsigned)nb, SOLVE_THREADS, smem, st>>>(b);
    d->timer.end(st);
    CK(cudaGetLastError());
    b.sh = nullptr;
    if (use_fast) {
      // tensor-core sign vote: E[p][b][l] = sum_i xh_ip * fp16(c_bi t_bil), fp32 accumulate, in chunks
      // of FAST_RC rows (cuBLAS: a plain fp16 GEMM, m = nb*L, n = Ppad, k = rows of the chunk)
      __half* BT = (__half*)(base + bb.BT);
      float* Cf = (float*)(base + bb.Cf);
      int nsl_pad = 1;
      while (nsl_pad * SLOT < h.kmax) nsl_pad <<= 1;
      // replicate lanes per CTA: SG_THREADS / (L * nsl_pad); with single-slot blocks the warp layout is
      // (8 LVs x 4 lanes) x ceil(L/8) LV groups, i.e. lanes come in fours
      int nbl_host = SG_THREADS / (h.L * nsl_pad);
      if (nsl_pad == 1 && nbl_host >= 4) nbl_host = (SG_THREADS / 32 / ((h.L + 7) / 8)) * 4;
      const int reps_per_cta = std::max(1, nbl_host) * SG_RPT;
      // rows per staged tile: at most 64, and small enough for two CTAs per SM
      const size_t sg_row_bytes = (size_t)h.Ppad * 4 + (size_t)reps_per_cta * 4;
      const int SG_ROWS = (int)std::max<size_t>(1, std::min<size_t>(SG_MAX_ROWS, (size_t)(d->max_smem / 2 - 8192) / sg_row_bytes));
      const size_t sg_smem = (size_t)SG_ROWS * sg_row_bytes;
      auto sg_kernel = (nsl_pad == 1) ? scoregen_kernel<true> : scoregen_kernel<false>;
      CK(cudaFuncSetAttribute(sg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sg_smem));
      const float one = 1.f, zero = 0.f;
      const int64_t ldl = (nb + 7) / 8 * 8;
      const int gemm_m = (int)(ldl * h.L);
      // padded replicate columns of the score buffer are only written up to the CTA grid's reach: clear once
      CK(cudaMemsetAsync(BT, 0, (size_t)gemm_m * FAST_RC * sizeof(__half), st));
      for (int64_t i0 = 0; i0 < d->N; i0 += FAST_RC) {
        const int rc = (int)std::min<int64_t>(FAST_RC, d->N - i0);
        const unsigned gy = (unsigned)((nb + reps_per_cta - 1) / reps_per_cta);
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((rc + SG_ROWS - 1) / SG_ROWS,
                                                                            (4 * d->sm_count + gy - 1) / gy));
        dim3 grid_sg(gx, gy);
        d->timer.begin(ST_SCOREGEN, st);
        sg_kernel<<<grid_sg, SG_THREADS, sg_smem, st>>>(d->Xf, counts_dev, D(bb.wf), D(bb.sh), d->N, h.Ppad, h.L,
                                                             m->dv.lv_off, m->dv.lv_k, nsl_pad, SG_ROWS, nb, ldl, i0, rc,
                                                             BT);
        d->timer.end(st);
        CK(cudaGetLastError());
        // C[nb*L x Ppad] (+)= B^T-free NT product: A = scores stored [nb*L x rc] (column-major view of the
        // row-major [rc][nb*L] buffer), B = xh chunk stored [Ppad x rc]
        d->timer.begin(ST_CROSS, st);
        cublasStatus_t cs = cublasGemmEx(d->blas, CUBLAS_OP_N, CUBLAS_OP_T, gemm_m, h.Ppad, rc, &one, BT, CUDA_R_16F,
                                         gemm_m, d->Xh + i0 * h.Ppad, CUDA_R_16F, h.Ppad, i0 == 0 ? &zero : &one, Cf,
                                         CUDA_R_32F, gemm_m, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
        d->timer.end(st);
        if (cs != CUBLAS_STATUS_SUCCESS) return fail(PLSPM_ERR_CUDA, "cublasGemmEx failed: " + std::to_string((int)cs));
      }
      b.fast_cross = Cf; b.inv_sd = d->inv_sd; b.fast_nb = ldl;
    } else {
      if (int rc = launch_stream(d, true, nb, counts_dev, bp.cross, D(bb.CG), D(bb.CGpart), D(bb.wf))) return rc;
    }
  }
  b.phase = h.full ? 0 : (use_fast ? 3 : 2);
  b.out_rows = out_rows; b.out_stride = h.n_out();
  if (single_fit) {
    b.weights = D(bb.weights); b.loadings = D(bb.loadings); b.r2 = D(bb.r2); b.paths = D(bb.paths);
    b.total = D(bb.totalfx); b.crossloadings = D(bb.crossl); b.score_coef = D(bb.coef); b.score_shift = D(bb.shift);
  }
  d->timer.begin(ST_SOLVE, st);
  solve_kernel<<<(unsigned)nb, SOLVE_THREADS, smem, st>>>(b);
  d->timer.end(st);
  CK(cudaGetLastError());
  return 0;
}

int plspm_fit(const plspm_model* m, const plspm_data* dc, int32_t scheme, double tol, int32_t max_iter, double* weights,
              double* loadings, double* r_squared, double* paths, double* total_effects, double* crossloadings,
              double* scores, int32_t* iters, int32_t* status) {
  if (!m || !dc || !same_layout(dc->model, m)) return fail(PLSPM_ERR_INVALID, "model/data mismatch");
  if (scheme < 0 || scheme > 2) return fail(PLSPM_ERR_INVALID, "unknown scheme");
  plspm_data* d = const_cast<plspm_data*>(dc);
  d->model = m;
  const HostModel& h = m->h;
  const int64_t N = d->N;
  const size_t L = h.L, P = h.P;
  BatchPlan bp;
  if (int rc = plan_batch(d, 1, bp)) return rc;
  const BatchBuffers bb = layout_batch(d, 1, bp, false, false, false, true, scores != nullptr);
  if (int rc = ws_reserve(d, bb.total)) return rc;
  char* base = (char*)d->ws.ptr;
  auto D = [&](size_t o) { return (double*)(base + o); };
  cudaStream_t st = d->stream;
  if (m->numeric) {
    if (crossloadings && !h.full)
      return fail(PLSPM_ERR_UNSUPPORTED, "numeric non-metric fit: crossloadings need a model with the full tile set");
    if (int rc = run_batch_num(d, 1, nullptr, bb, scheme, tol, max_iter, bp, nullptr, true)) return rc;
  } else if (int rc = run_batch(d, 1, nullptr, bb, scheme, tol, max_iter, bp, nullptr, true)) {
    return rc;
  }
  if (scores) {
    d->timer.begin(ST_SCORES, st);
    scores_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N, h.Ppad, h.L, m->dv.lv_off, m->dv.lv_k, D(bb.coef),
                                                   D(bb.shift), D(bb.scores));
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  int host_it = 0, host_st = 0;
  CK(cudaMemcpyAsync(&host_it, base + bb.iters, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&host_st, base + bb.status, 4, cudaMemcpyDeviceToHost, st));
  if (weights) CK(cudaMemcpyAsync(weights, D(bb.weights), P * 8, cudaMemcpyDeviceToHost, st));
  if (loadings) CK(cudaMemcpyAsync(loadings, D(bb.loadings), P * 8, cudaMemcpyDeviceToHost, st));
  if (r_squared) CK(cudaMemcpyAsync(r_squared, D(bb.r2), L * 8, cudaMemcpyDeviceToHost, st));
  if (paths) CK(cudaMemcpyAsync(paths, D(bb.paths), L * L * 8, cudaMemcpyDeviceToHost, st));
  if (total_effects) CK(cudaMemcpyAsync(total_effects, D(bb.totalfx), L * L * 8, cudaMemcpyDeviceToHost, st));
  if (crossloadings) CK(cudaMemcpyAsync(crossloadings, D(bb.crossl), P * L * 8, cudaMemcpyDeviceToHost, st));
  if (scores) CK(cudaMemcpyAsync(scores, D(bb.scores), (size_t)N * L * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  d->timer.collect();
  if (iters) *iters = host_it;
  if (status) *status = host_st;
  return PLSPM_OK;
}

int plspm_bootstrap(const plspm_model* m, const plspm_data* dc, int32_t scheme, double tol, int32_t max_iter,
                    int64_t rep_begin, int64_t rep_count, uint64_t seed, const int32_t* idx, double* out,
                    int32_t out_is_device, int32_t* status, int32_t* iters) {
  if (!m || !dc || !same_layout(dc->model, m)) return fail(PLSPM_ERR_INVALID, "model/data mismatch");
  if (scheme < 0 || scheme > 2) return fail(PLSPM_ERR_INVALID, "unknown scheme");
  if (rep_count < 0 || !out) return fail(PLSPM_ERR_INVALID, "bad replicate range / null output");
  if (rep_count == 0) return PLSPM_OK;
  plspm_data* d = const_cast<plspm_data*>(dc);
  d->model = m;
  const HostModel& h = m->h;
  const int64_t N = d->N;
  const size_t n_out = h.n_out();
  if (idx)
    for (int64_t e = 0; e < rep_count * N; ++e)
      if (idx[e] < 0 || idx[e] >= N) return fail(PLSPM_ERR_INVALID, "resample index out of range");
  // batch size: bound the workspace (~1.5 GB) and keep the Gram grid a whole number of waves
  const bool vote = !h.full && !m->numeric;
  const size_t per_rep = (size_t)N * 4 + ((size_t)h.n_tiles * TILE + (vote ? (size_t)h.n_cross * TILE : 0) +
                                          (m->numeric ? 6 : 2) * h.Ppad + h.ws_doubles + n_out) * 8 +
                         (idx ? (size_t)N * 4 : 0) + (d->i8_colsum ? (size_t)d->Npad + I8_DIGITS * ((size_t)h.Ppad + d->n_zcols) * 4 : 0) + 64;
  // the planes of a chunk are regenerated for every batch in streaming mode: large batches amortise that
  const size_t ws_budget = d->n_zcols && !d->Z8 ? (size_t)12 << 30 : (size_t)1536 << 20;
  int64_t nb_max = std::max<int64_t>(1, (int64_t)(ws_budget / per_rep));
  if (getenv("PLSPM_MAX_BATCH")) nb_max = std::max<int64_t>(1, std::min<int64_t>(nb_max, atoll(getenv("PLSPM_MAX_BATCH"))));
  nb_max = std::min<int64_t>(nb_max, rep_count);
  const int64_t wave = (int64_t)d->sm_count * GRAM_WARPS;  // warp items per wave
  const int64_t items_per_rep = vote ? std::max(h.n_tg, h.n_tg_cross) : h.n_tg;
  if (nb_max * items_per_rep > wave) {
    int64_t waves = nb_max * items_per_rep / wave;
    nb_max = std::max<int64_t>(1, waves * wave / items_per_rep);
  }
  BatchPlan bp;
  if (int rc = plan_batch(d, nb_max, bp)) return rc;
  const BatchBuffers bb = layout_batch(d, nb_max, bp, true, idx != nullptr, out_is_device != 0, false, false);
  if (int rc = ws_reserve(d, bb.total)) return rc;
  char* base = (char*)d->ws.ptr;
  cudaStream_t st = d->stream;
  CK(cudaMemsetAsync(base + bb.ovf, 0, 8, st));
  for (int64_t b0 = 0; b0 < rep_count; b0 += nb_max) {
    const int64_t nb = std::min(nb_max, rep_count - b0);
    // a short last batch reuses the plan (and therefore the workspace layout) of a full one
    uint32_t* cnt = (uint32_t*)(base + bb.counts);
    int32_t* idx_dev = nullptr;
    CK(cudaMemsetAsync(cnt, 0, (size_t)nb * N * 4, st));
    if (idx) {
      idx_dev = (int32_t*)(base + bb.idx);
      CK(cudaMemcpyAsync(idx_dev, idx + b0 * N, (size_t)nb * N * 4, cudaMemcpyHostToDevice, st));
    }
    const int64_t threads = ((N + 3) / 4) * nb;
    d->timer.begin(ST_COUNTS, st);
    counts_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(cnt, idx_dev, N, nb, rep_begin + b0, seed);
    d->timer.end(st);
    CK(cudaGetLastError());
    double* rows = out_is_device ? out + (size_t)b0 * n_out : (double*)(base + bb.out);
    for (int attempt = 0; attempt < 2; ++attempt) {
      if (m->numeric) {
        if (int rc = run_batch_num(d, nb, cnt, bb, scheme, tol, max_iter, bp, rows, false)) return rc;
      } else if (int rc = run_batch(d, nb, cnt, bb, scheme, tol, max_iter, bp, rows, false)) {
        return rc;
      }
      if (!d->i8_colsum) break;
      // a multiplicity above 127 does not fit the int8 operand of the tensor-core column sums: redo in fp64
      int ovf = 0;
      CK(cudaMemcpyAsync(&ovf, base + bb.ovf, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (!ovf) break;
      CK(cudaMemsetAsync(base + bb.ovf, 0, 8, st));
      d->i8_colsum = false;
    }
    if (vote && d->fast_vote) {
      // replicates whose low-precision sign vote was undecided are redone with exact fp64 cross moments
      std::vector<int> st_host(nb);
      CK(cudaMemcpyAsync(st_host.data(), base + bb.status, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      std::vector<int> redo;
      for (int64_t r = 0; r < nb; ++r)
        if (st_host[r] == STATUS_AMBIGUOUS) redo.push_back((int)r);
      if (!redo.empty()) {
        int* map_dev = (int*)(base + bb.rep_map);
        CK(cudaMemcpyAsync(map_dev, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, st));
        if (int rc = redo_exact(d, (int64_t)redo.size(), map_dev, cnt, bb, scheme, tol, max_iter, bp, rows)) return rc;
        g_redo_count += (int64_t)redo.size();
        if ((int64_t)redo.size() * 2 > nb) d->fast_vote = false;  // this data does not suit the fp16 vote
      }
    }
    if (!out_is_device)
      CK(cudaMemcpyAsync(out + (size_t)b0 * n_out, rows, (size_t)nb * n_out * 8, cudaMemcpyDeviceToHost, st));
    if (iters) CK(cudaMemcpyAsync(iters + b0, base + bb.iters, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
    if (status) CK(cudaMemcpyAsync(status + b0, base + bb.status, (size_t)nb * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    d->timer.collect();
  }
  return PLSPM_OK;
}

int plspm_bootstrap_host(const plspm_model* m, const double* X, int64_t N, int64_t ld, int32_t scheme, double tol,
                         int32_t max_iter, int64_t rep_begin, int64_t rep_count, uint64_t seed, const int32_t* idx,
                         double* out, int32_t* status, int32_t* iters) {
  plspm_data* d = nullptr;
  int rc = plspm_data_create(m, X, N, ld, 0, &d);
  if (rc) return rc;
  rc = plspm_bootstrap(m, d, scheme, tol, max_iter, rep_begin, rep_count, seed, idx, out, 0, status, iters);
  plspm_data_destroy(d);
  return rc;
}

int plspm_resample_indices(uint64_t seed, int64_t replicate, int64_t N, int32_t* idx_out) {
  if (!idx_out || N < 1 || N >= ((int64_t)1 << 32)) return fail(PLSPM_ERR_INVALID, "bad arguments");
  int32_t* dev = nullptr;
  CK(cudaMalloc((void**)&dev, (size_t)N * 4));
  const int64_t groups = (N + 3) / 4;
  indices_kernel<<<(unsigned)((groups + 255) / 256), 256>>>(dev, N, (uint64_t)replicate, seed);
  cudaError_t e = cudaMemcpy(idx_out, dev, (size_t)N * 4, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) return fail(PLSPM_ERR_CUDA, cudaGetErrorString(e));
  return PLSPM_OK;
}

int plspm_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (int i = 0; i < ST_N; ++i) { g_prof.ms[i] = 0; g_prof.launches[i] = 0; }
  return PLSPM_OK;
}
int plspm_profile_get(double* ms, int64_t* launches) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (int i = 0; i < ST_N; ++i) {
    if (ms) ms[i] = g_prof.ms[i];
    if (launches) launches[i] = g_prof.launches[i];
  }
  return PLSPM_OK;
}

int plspm_redo_count(int64_t* count) {
  if (!count) return fail(PLSPM_ERR_INVALID, "null argument");
  *count = g_redo_count;
  return PLSPM_OK;
}

int plspm_host_alloc(void** ptr, int64_t bytes) {
  CK(cudaMallocHost(ptr, (size_t)bytes));
  return PLSPM_OK;
}
int plspm_host_free(void* ptr) {
  CK(cudaFreeHost(ptr));
  return PLSPM_OK;
}

}  // extern "C"
