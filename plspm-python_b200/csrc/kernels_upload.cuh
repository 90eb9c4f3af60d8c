// Upload: column means, slot relayout + centring, fp16 / fp32 copies.
// Part of the single translation unit plspm_b200.cu (included there, in this order); see DESIGN.md §4.
#pragma once

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
// Second stage of the two-stage column reductions: ONE WARP per column (launch <<<(P + 3) / 4, 128>>>).  Lane l owns
// the partials l, l + 32, ... as four independent chains, then a xor butterfly -- a fixed order, so the result is
// deterministic and identical in every lane.  (One thread per column walking all 1024 partials was a 100 us chain of
// dependent loads per kernel, five kernels per upload: 0.5 ms of every end-to-end call.)
template <typename T, typename Op>
__device__ __forceinline__ T warp_column_reduce(const T* __restrict__ partial, int nblocks, int64_t stride, int p, T init, Op op) {
  const int lane = threadIdx.x & 31;
  T a0 = init, a1 = init, a2 = init, a3 = init;
  int b = lane;
  for (; b + 96 < nblocks; b += 128) {
    a0 = op(a0, partial[(int64_t)b * stride + p]);
    a1 = op(a1, partial[(int64_t)(b + 32) * stride + p]);
    a2 = op(a2, partial[(int64_t)(b + 64) * stride + p]);
    a3 = op(a3, partial[(int64_t)(b + 96) * stride + p]);
  }
  for (; b < nblocks; b += 32) a0 = op(a0, partial[(int64_t)b * stride + p]);
  T s = op(op(a0, a1), op(a2, a3));
  for (int o = 16; o; o >>= 1) s = op(s, __shfl_xor_sync(0xffffffffu, s, o));
  return s;
}
struct OpAdd { template <typename T> __device__ T operator()(T a, T b) const { return a + b; } };
struct OpMax { __device__ double operator()(double a, double b) const { return fmax(a, b); } };

__global__ void colmean_final_kernel(const double* __restrict__ partial, int nblocks, int P, int64_t N,
                                     const int* __restrict__ src_col, double* __restrict__ mu) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const double s = warp_column_reduce(partial, nblocks, (int64_t)P, p, 0.0, OpAdd());
  if ((threadIdx.x & 31) == 0) mu[src_col[p]] = s / (double)N;
}
// plain column totals (the sums of the centred columns kept in the data handle)
__global__ void colsum_final_kernel(const double* __restrict__ partial, int nblocks, int P, double* __restrict__ out) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= P) return;
  const double s = warp_column_reduce(partial, nblocks, (int64_t)P, p, 0.0, OpAdd());
  if ((threadIdx.x & 31) == 0) out[p] = s;
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
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (p >= Ppad) return;
  const double s = warp_column_reduce(partial, nblocks, (int64_t)Ppad, p, 0.0, OpAdd());
  if ((threadIdx.x & 31) == 0) inv_sd[p] = s > 0.0 ? 1.0 / sqrt(s / (double)N) : 0.0;
}
__global__ void make_half_kernel(const double* __restrict__ X, int64_t N, int Ppad, const double* __restrict__ inv_sd,
                                 __half* __restrict__ out) {
  const int64_t total = N * Ppad;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x)
    out[e] = __double2half(X[e] * inv_sd[e % Ppad]);
}
