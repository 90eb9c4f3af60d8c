// Streaming passes over the scores: sign-vote score generation, non-metric stopping criterion.
// Part of the single translation unit plspm_b200.cu (included there, in this order); see DESIGN.md §4.
#pragma once

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
