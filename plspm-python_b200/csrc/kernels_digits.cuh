// Integer digit planes: exact tensor-core column sums and Gram tiles.
// Part of the single translation unit plspm_b200.cu (included there, in this order); see DESIGN.md §4.
#pragma once

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
