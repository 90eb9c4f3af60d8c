// Bootstrap on data with missing values: the reference re-imputes every replicate with the column means of ITS
// observed rows (bootstrap.py:57 -> estimator.py:33 -> config.py:299-305 -> util.py:61-68).  In the moment domain that
// is a closed form over an AUGMENTED matrix [x0 | m]: x0 = the data with missing entries set to 0, m_p = the missing
// indicator of column p (only for columns that have missing entries).  With the replicate's mean of the observed
// entries  mu_p = sum_i c_i x0_ip / sum_i c_i (1 - m_ip)  the imputed value is x^_ip = x0_ip + m_ip mu_p, so every
// first and second moment of the imputed replicate is a polynomial in mu and the moments of [x0 | m]:
//   sum c x^_p x^_q = sum c x0_p x0_q + mu_q sum c x0_p m_q + mu_p sum c m_p x0_q + mu_p mu_q sum c m_p m_q
// The engine computes the moments of the augmented matrix for the whole batch with its usual kernels (tcgen05
// integer Gram or fp64), this kernel turns them into the moments of the imputed data in the BASE model's tile
// layout, and the solver runs on those.  Part of the single translation unit plspm_b200.cu.
#pragma once

// Data are held centred at upload: a = x0 - alpha, e = m - eps (alpha, eps = upload means, mu_a[] below), and the
// base moments are taken about kappa_p = alpha_p:  x~_p = x^_p - alpha_p = a_p + mu_p (e_p + eps_p).
// CTA = one replicate.  shared: mu | eps | Sa | Se, each [base Ppad].
__global__ void __launch_bounds__(256) impute_moments_kernel(const plspm::ModelView Ma, const plspm::ModelView Mb,
                                                             const double* __restrict__ Ga, int64_t ga_stride,
                                                             const double* __restrict__ csa, int64_t csa_stride,
                                                             const double* __restrict__ mu_a, const int* __restrict__ ax,
                                                             const int* __restrict__ am, double N, double* __restrict__ Gb,
                                                             int64_t gb_stride, double* __restrict__ csb, int64_t csb_stride) {
  extern __shared__ double im_sm[];
  const int Pb = Mb.Ppad;
  double *mu = im_sm, *eps = mu + Pb, *Sa = eps + Pb, *Se = Sa + Pb;
  const int64_t b = blockIdx.x;
  const double* G = Ga + b * ga_stride;
  const double* cs = csa + b * csa_stride;
  for (int p = threadIdx.x; p < Pb; p += blockDim.x) {
    const int a = ax[p], e = am[p];
    double sa = 0.0, se = 0.0, ep = 0.0, m = 0.0;
    if (a >= 0) {
      sa = cs[a];
      if (e >= 0) {
        se = cs[e];
        ep = mu_a[e];
        m = (sa + N * mu_a[a]) / (N - se - N * ep);  // (a column with no observed row in the replicate: NaN, like the reference)
      }
    }
    mu[p] = m; eps[p] = ep; Sa[p] = sa; Se[p] = se;
    csb[b * csb_stride + p] = sa + m * (se + ep * N);
  }
  __syncthreads();
  double* out = Gb + b * gb_stride;
  for (int t = threadIdx.x; t < Mb.n_tiles * TILE; t += blockDim.x) {
    const int tile = t / TILE, rc = t - tile * TILE, r = rc / SLOT, c = rc - r * SLOT;
    const int p = Mb.tile_sa[tile] * SLOT + r, q = Mb.tile_sb[tile] * SLOT + c;
    const int ap = ax[p], aq = ax[q];
    double v = 0.0;
    if (ap >= 0 && aq >= 0) {
      const int ep = am[p], eq = am[q];
      v = plspm::gram_raw(Ma, G, ap, aq);
      if (eq >= 0) v += mu[q] * (plspm::gram_raw(Ma, G, ap, eq) + eps[q] * Sa[p]);
      if (ep >= 0) v += mu[p] * (plspm::gram_raw(Ma, G, ep, aq) + eps[p] * Sa[q]);
      if (ep >= 0 && eq >= 0)
        v += mu[p] * mu[q] * (plspm::gram_raw(Ma, G, ep, eq) + eps[q] * Se[p] + eps[p] * Se[q] + eps[p] * eps[q] * N);
    }
    out[t] = v;
  }
}

__global__ void gather_kernel(const double* __restrict__ src, const int* __restrict__ map, int n, double* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = map[i] >= 0 ? src[map[i]] : 0.0;
}
