// Non-metric PLS-PM with numeric scales (Scale.NUM / Scale.RAW, complete data) in the covariance
// domain: reference `_NonmetricWeights` (weights.py:73-133), Mode A / B (mode.py:31-42, 54-61),
// treatment config.py:306-319.  SURVEY.md §8(f) row f3.
//
// Differences from the metric solver (solver_core.h):
//   * every manifest variable is standardised to unit POPULATION variance within the (re)sample, so
//     S is the correlation matrix:  S_pq = (G_pq/N - m_p m_q) / (sd_p sd_q);
//   * the state is the coefficient vector a_l of every LV score, Y_l = Xstd_l a_l, re-normalised to unit
//     variance after every outer step (treat_numpy(Y) * correction, mode.py:41); initial a_l = 1/sqrt(K_l);
//   * no sign vote; the reported weights are a_l (weights.py:128-131);
//   * the stopping rule is on the SCORES, sum_{i,l} (|y_old,il| - |y_new,il|)^2 < tol (weights.py:120).  That
//     sum is not a function of second moments, so the iteration is driven from the host: one call of
//     num_step() per outer iteration for all replicates of a batch, followed by one streaming pass
//     (conv_kernel) that evaluates the criterion from the old and new coefficient vectors.
//     Since (|a| - |b|)^2 = (a - b)^2 + 4ab [ab < 0], the criterion splits into
//        N sum_l (a_old - a_new)_l' S_ll (a_old - a_new)_l      -- second moments, computed here (conv_main)
//      + 4 sum_{i,l : y_old y_new < 0} c_i y_old,il y_new,il    -- only the rows whose score changes sign,
//     so the streaming pass only has to find the sign changes (scores close to zero).
//
// Same dual compilation as solver_core.h (device: one CTA per replicate; host: emulation for tests).
#pragma once
#include "solver_core.h"

namespace plspm {

struct NumStepArgs {
  ModelView M;
  const double* G;       // [n_tiles*64]
  const double* colsum;  // [Ppad]
  double N;
  int scheme;
  double tol;
  int max_iter;
  double conv_in;        // criterion of the previous step (ignored while it == 0)
  double* conv_main;     // [1] out: the second-moment part of this step's criterion (see below)
  double* ws;            // [M.ws_doubles] global scratch private to this replicate (persists across steps)
  // persistent state of the replicate (global memory)
  double* a;             // [Ppad] current coefficient vectors
  int* meta;             // [4]: 0 iterate() calls made, 1 done, 2 status
  // for the criterion pass: y = x~ . coef - shift, before and after this step
  double *coef_old, *coef_new;    // [Ppad]
  double *shift_old, *shift_new;  // [L]
  // outputs, written when the replicate finishes (any may be null)
  double* out_row;
  double *weights, *loadings, *r2, *paths, *total, *crossloadings, *score_coef, *score_shift;
  int *iters, *status;
};

// Shared memory: same carve-up as solve_replicate (HostModel::solver_smem_doubles()).
PL_HD void num_step(const NumStepArgs& A, double* smem) {
  const ModelView& M = A.M;
  const int L = M.L, Ppad = M.Ppad, tid = PL_TID, nt = PL_NT;
  double* a = smem;                  // [Ppad] coefficients (current)
  double* an = a + Ppad;             // [Ppad] next
  double* isd = an + Ppad;           // [Ppad] 1/sd of every column in this (re)sample
  double* m = isd + Ppad;            // [Ppad] column means of x~
  double* V = m + Ppad;              // [n_v]
  double* var = V + M.n_v;           // [L]
  double* r2 = var + L;              // [L]
  double* tmp = r2 + 2 * L;          // (two spare [L] arrays of the layout are skipped)
  double* R = tmp;                   // [L*L] covariance (later correlation) of the scores
  double* E = R + (size_t)L * L;     // [L*L]
  double* Bm = E + (size_t)L * L;    // [L*L]
  double* red = Bm + (size_t)L * L;  // [40]
  int* flag = (int*)(red + 34);
  const double N = A.N, invN = 1.0 / A.N;

  if (A.meta[1]) return;             // already finished (uniform: every thread reads the same word)
  const int it = A.meta[0];
  for (int p = tid; p < Ppad; p += nt) {
    m[p] = A.colsum[p] * invN;
    double v = (M.col_lv[p] >= 0) ? gram_raw(M, A.G, p, p) * invN - m[p] * m[p] : 0.0;
    isd[p] = v > 0.0 ? 1.0 / sqrt(v) : 0.0;
    a[p] = (it == 0) ? ((M.col_lv[p] >= 0) ? 1.0 / sqrt((double)M.lv_k[M.col_lv[p]]) : 0.0) : A.a[p];
  }
  if (tid == 0) flag[0] = A.meta[2];
  PL_SYNC();
#define PLN_S(p, q) ((gram_raw(M, A.G, (p), (q)) * invN - m[p] * m[q]) * isd[p] * isd[q])

  const bool stop = (it >= 1) && ((A.conv_in < A.tol) || (it > A.max_iter) || flag[0] != STATUS_OK);
  const int ols_stride = ols_scratch_doubles(M.max_deg);
  double* ols = A.ws + (M.ws_doubles - L * ols_stride);

  // ---- V_d = S_lj a_j, variances and covariances of the current scores --------------------------
  for (int t = tid; t < M.n_pairs * M.kmax; t += nt) {
    int d = t / M.kmax, r = t - d * M.kmax;
    int l = M.pair_l[d], j = M.pair_j[d];
    if (r >= M.lv_k[l]) continue;
    int kj = M.lv_k[j], ol = M.lv_off[l], oj = M.lv_off[j];
    double acc = 0.0;
    for (int c = 0; c < kj; ++c) acc += PLN_S(ol + r, oj + c) * a[oj + c];
    V[M.pair_voff[d] + r] = acc;
  }
  PL_SYNC();
  for (int d = tid; d < M.n_pairs; d += nt) {
    int l = M.pair_l[d], j = M.pair_j[d];
    if (l >= j) {
      int o = M.lv_off[l];
      double acc = 0.0;
      for (int r = 0; r < M.lv_k[l]; ++r) acc += a[o + r] * V[M.pair_voff[d] + r];
      R[l * L + j] = acc;
      R[j * L + l] = acc;
      if (l == j) var[l] = acc;
    }
  }
  PL_SYNC();

  if (stop) {
    // ---- finish: weights = a, inner model on the (unit variance) scores, loadings -----------------
    int status = flag[0];
    if (status == STATUS_OK && it > A.max_iter) status = STATUS_NOT_CONVERGED;  // weights.py:185 (Q4)
    for (int e = tid; e < L * L; e += nt) {
      int i = e / L, j = e - i * L;
      R[e] = R[e] / sqrt(var[i] * var[j]);
      Bm[e] = 0.0;
    }
    PL_SYNC();
    for (int i = tid; i < L; i += nt) {
      int n = M.pred_begin[i + 1] - M.pred_begin[i];
      double rr = 0.0;
      if (n > 0) {
        double* sc = ols + (size_t)i * ols_stride;
        double* beta = sc + M.max_deg * M.max_deg;
        if (!regress_on_predecessors(M, R, i, sc, beta)) {
          flag[0] = STATUS_SINGULAR;
        } else {
          for (int k = 0; k < n; ++k) {
            int j = M.pred_idx[M.pred_begin[i] + k];
            Bm[i * L + j] = beta[k];
            rr += beta[k] * R[j * L + i];
          }
        }
      }
      r2[i] = rr;
    }
    PL_SYNC();
    if (status == STATUS_OK) status = flag[0];
    double* T = E;
    for (int j = tid; j < L; j += nt)
      for (int i = 0; i < L; ++i) {
        double acc = Bm[i * L + j];
        for (int k = j + 1; k < i; ++k) acc += Bm[i * L + k] * T[k * L + j];
        T[i * L + j] = (i > j) ? acc : 0.0;
      }
    PL_SYNC();
    const int P = M.P, ne = M.n_eff;
    for (int p = tid; p < Ppad; p += nt) {
      int l = M.col_lv[p];
      if (l < 0) {
        if (A.score_coef) A.score_coef[p] = 0.0;
        continue;
      }
      int src = M.col_src[p];
      double load = V[M.pair_voff[M.lv_pair_begin[l]] + (p - M.lv_off[l])] / sqrt(var[l]);  // corr(x_p, y_l)
      if (A.out_row) {
        A.out_row[src] = a[p];
        A.out_row[P + L + 2 * ne + src] = load;
      }
      if (A.weights) A.weights[src] = a[p];
      if (A.loadings) A.loadings[src] = load;
      if (A.score_coef) A.score_coef[p] = a[p] * isd[p];
    }
    for (int l = tid; l < L; l += nt) {
      if (A.out_row) A.out_row[P + l] = r2[l];
      if (A.r2) A.r2[l] = r2[l];
      if (A.score_shift) {
        double sh = 0.0;
        for (int r = 0; r < M.lv_k[l]; ++r) sh += m[M.lv_off[l] + r] * a[M.lv_off[l] + r] * isd[M.lv_off[l] + r];
        A.score_shift[l] = sh;
      }
    }
    for (int e = tid; e < ne; e += nt) {
      int f = M.eff_from[e], t = M.eff_to[e];
      if (A.out_row) {
        A.out_row[P + L + e] = T[t * L + f];
        A.out_row[P + L + ne + e] = Bm[t * L + f];
      }
    }
    for (int e = tid; e < L * L; e += nt) {
      if (A.paths) A.paths[e] = Bm[e];
      if (A.total) A.total[e] = T[e];
    }
    if (A.crossloadings && M.full)  // corr(x_p, y_l) for every pair needs the full tile set
      for (int t = tid; t < Ppad * L; t += nt) {
        int p = t / L, l = t - p * L;
        if (M.col_lv[p] < 0) continue;
        int o = M.lv_off[l];
        double acc = 0.0;
        for (int c = 0; c < M.lv_k[l]; ++c) acc += PLN_S(p, o + c) * a[o + c];
        A.crossloadings[(size_t)M.col_src[p] * L + l] = acc / sqrt(var[l]);
      }
    if (tid == 0) {
      A.meta[1] = 1;
      A.meta[2] = status;
      if (A.iters) *A.iters = it;
      if (A.status) *A.status = status;
    }
    return;
  }

  // ---- one outer iteration (weights.py:107-120) ---------------------------------------------------
  if (it == 0)  // Mode B: factor the block correlation matrix once per replicate
    for (int l = tid; l < L; l += nt)
      if (M.lv_mode[l] == MODE_B) {
        int o = M.lv_off[l], k = M.lv_k[l];
        double* C = A.ws + M.chol_b_off[l];
#define PLN_BLK(r_, c_) PLN_S(o + (r_), o + (c_))
        PL_MODE_B_PREPARE(C, k, PLN_BLK);
#undef PLN_BLK
      }
  // inner weights from the covariance of the current scores (scheme.py)
  if (A.scheme == SCHEME_PATH) {
    for (int e = tid; e < L * L; e += nt) E[e] = 0.0;
    PL_SYNC();
    for (int i = tid; i < L; i += nt) {
      int n = M.pred_begin[i + 1] - M.pred_begin[i];
      if (n > 0) {  // OLS without intercept of Y_i on its predecessors: Cov_pp beta = Cov_pi
        double* sc = ols + (size_t)i * ols_stride;
        double* beta = sc + M.max_deg * M.max_deg;
        if (!regress_on_predecessors(M, R, i, sc, beta)) flag[0] = STATUS_SINGULAR;
        for (int k = 0; k < n; ++k) E[M.pred_idx[M.pred_begin[i] + k] * L + i] = beta[k];
      }
      for (int k = M.succ_begin[i]; k < M.succ_begin[i + 1]; ++k) {
        int s = M.succ_idx[k];
        E[s * L + i] = R[s * L + i] / sqrt(var[s] * var[i]);
      }
    }
  } else {
    for (int e = tid; e < L * L; e += nt) {
      int j = e / L, l = e - j * L;
      double val = 0.0;
      if (M.path[j * L + l] | M.path[l * L + j]) {
        double c = R[e];
        if (A.scheme == SCHEME_CENTROID) val = (c > 0.0) ? 1.0 : ((c < 0.0) ? -1.0 : c);  // sign of the correlation
        else val = c * N / (N - 1.0);  // np.cov, ddof = 1 (scheme.py:37)
      }
      E[e] = val;
    }
  }
  PL_SYNC();
  // outer step: cov(X_l, Z_l) = sum_j E[j,l] S_lj a_j ; Mode B solves with the block correlation
  for (int p = tid; p < Ppad; p += nt) {
    int l = M.col_lv[p];
    if (l < 0) { an[p] = 0.0; continue; }
    int r = p - M.lv_off[l];
    double acc = 0.0;
    for (int d = M.lv_pair_begin[l] + 1; d < M.lv_pair_begin[l + 1]; ++d) {
      double e = E[M.pair_j[d] * L + l];
      if (e != 0.0) acc += e * V[M.pair_voff[d] + r];
    }
    an[p] = acc;
  }
  PL_SYNC();
  for (int l = tid; l < L; l += nt) {
    int o = M.lv_off[l], k = M.lv_k[l];
    if (M.lv_mode[l] == MODE_B && flag[0] == STATUS_OK) mode_b_solve(A.ws + M.chol_b_off[l], k, an + o);
    double q = 0.0;  // variance of X_l w: w' S_ll w
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) q += an[o + r] * PLN_S(o + r, o + c) * an[o + c];
    double s = 1.0 / sqrt(q);
    for (int r = 0; r < k; ++r) an[o + r] *= s;  // treat_numpy(Y) * correction: unit population variance
  }
  PL_SYNC();
  // second-moment part of the stopping criterion: N (a - an)_l' S_ll (a - an)_l summed over the LVs
  for (int l = tid; l < L; l += nt) {
    int o = M.lv_off[l], k = M.lv_k[l];
    double q = 0.0;
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) q += (a[o + r] - an[o + r]) * PLN_S(o + r, o + c) * (a[o + c] - an[o + c]);
    r2[l] = q;
  }
  PL_SYNC();
  if (tid == 0 && A.conv_main) {
    double q = 0.0;
    for (int l = 0; l < L; ++l) q += r2[l];
    A.conv_main[0] = N * q;
  }
  // hand the old / new scores to the criterion pass and persist the state
  for (int p = tid; p < Ppad; p += nt) {
    A.coef_old[p] = a[p] * isd[p];
    A.coef_new[p] = an[p] * isd[p];
    A.a[p] = an[p];
  }
  for (int l = tid; l < L; l += nt) {
    double so = 0.0, sn = 0.0;
    for (int r = 0; r < M.lv_k[l]; ++r) {
      int p = M.lv_off[l] + r;
      so += m[p] * a[p] * isd[p];
      sn += m[p] * an[p] * isd[p];
    }
    A.shift_old[l] = so;
    A.shift_new[l] = sn;
  }
  if (tid == 0) {
    A.meta[0] = it + 1;
    A.meta[2] = flag[0];
  }
#undef PLN_S
}

}  // namespace plspm
