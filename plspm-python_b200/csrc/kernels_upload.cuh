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
