// Per-replicate PLS-PM solver in the covariance domain.
//
// One CTA solves one fit (the original sample or one bootstrap replicate) from the
// replicate's weighted second moments
//     G[p,q]   = sum_i c_i x~_ip x~_iq        (8x8 tiles, see plspm_model.h)
//     colsum_p = sum_i c_i x~_ip
// where x~ = x - mu (global column means removed once at upload) and c_i is the
// multiplicity of row i in the resample.  With m = colsum/N and
//     S = (G/N - m m') / scale^2       (population covariance of the TREATED data,
//                                       config.py:299-305, util.py:33-39)
// every quantity of the reference's iteration is a function of S:
//     Y = Xc W            ->  var(Y_l) = w_l' S_ll w_l,   R = D W' S W D
//     Z = Ystd E,  X_l' Z_l / N  ->  sum_j a d_j E[j,l] S_lj w_j           (mode.py:28-29)
//     lstsq(X_l, Z_l)            ->  S_ll^-1 (same vector)                  (mode.py:50-52)
// so the N-length arrays Y and Z are never formed (the Gram form
// reproduces the reference to 7e-16 in the same number of iterations).
//
// The same source compiles (a) as device code, one CTA per replicate, and (b) with
// PLSPM_HOST_EMUL as a single-"thread" host function that tests/ uses to check the solver
// logic against the oracle without a GPU.  (b) is test infrastructure, never shipped.
#pragma once
#include <math.h>

#include "plspm_model.h"

#if defined(__CUDACC__) && !defined(PLSPM_HOST_EMUL)
#define PL_DEVICE_BUILD 1
#define PL_HD __device__ __forceinline__
#define PL_TID ((int)threadIdx.x)
#define PL_NT ((int)blockDim.x)
#define PL_SYNC() __syncthreads()
#else
#define PL_DEVICE_BUILD 0
#define PL_HD inline
#define PL_TID 0
#define PL_NT 1
#define PL_SYNC() ((void)0)
#endif

namespace plspm {

struct SolveArgs {
  ModelView M;
  const double* G;       // [n_tiles*64] raw weighted Gram tiles of this replicate
  const double* colsum;  // [Ppad]
  const double* mu;      // [Ppad] global column means removed at upload
  double N;              // observations in the (re)sample
  int scheme;
  double tol;
  int max_iter;
  int phase;             // 0: everything (full tile set); 1: stop after the final weights (writes wf_out);
                         // 2: everything, cov(x_p, score_l) taken from the cross-moment tiles;
                         // 3: like 2, but LV pairs outside the tile set vote with the SIGN of a
                         //    low-precision (fp16 tensor-core) cross moment when it is provably right,
                         //    otherwise the replicate is handed back as STATUS_AMBIGUOUS
  const float* fast_cross;  // [Ppad][L][fast_nb] fp32 sums of xh_ip * (c_i t_il) (phase 3)
  int fast_uncentred;       // phase 3: the scores behind fast_cross were NOT centred (fused tcgen05 vote kernel,
                            // fp16 score MMA): remove sh_l * sum_i c_i xh_ip here and use that kernel's error terms
  const double* inv_sd;     // [Ppad] global 1/sd used to scale xh (phase 3)
  int64_t fast_nb, fast_b;  // replicate stride (batch size rounded up to 8) and this replicate's position
  double* sh_out;           // [L] score means sum_q m_q wf_q (phase 1)
  const double* cross;   // [n_cross*64] raw sum_i c_i x~_ip (x~_i . wf_l) tiles (phase 2)
  double* wf_out;        // [Ppad] final normalised weights, padded layout (phase 1)
  double* ws;            // [M.ws_doubles] global scratch private to this CTA
  double* state;         // [solver_state_doubles(M)] optional: phase 1 stores the converged iteration state here ...
  int resume;            // ... and phases 2 / 3 with resume != 0 continue from it instead of iterating again
  // outputs; any pointer may be null
  double* out_row;        // [2P+L+2n_eff] weights | r_squared | total effects | direct effects | loadings
  double* weights;        // [P]
  double* loadings;       // [P]
  double* r2;             // [L]
  double* paths;          // [L*L]
  double* total;          // [L*L]
  double* crossloadings;  // [P*L]
  double* score_coef;     // [Ppad] sign_l * wf_p / scale  (scores = (x~ - m) . coef)
  double* score_shift;    // [L]    sum_p m_p coef_p
  int* iters;
  int* status;
};

// iteration state handed from phase 1 to phases 2 / 3: w | V | dinv | R | {1/scale^2, iterations, status}
inline size_t solver_state_doubles(const ModelView& M) { return (size_t)M.Ppad + M.n_v + M.L + (size_t)M.L * M.L + 4; }

PL_HD double block_sum(double v, double* red) {
#if PL_DEVICE_BUILD
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  int nw = (blockDim.x + 31) >> 5;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
#else
  (void)red;
  return v;
#endif
}

PL_HD void vote_add(int* votes, int l, int v) {
#if PL_DEVICE_BUILD
  atomicAdd(&votes[l], v);
#else
  votes[l] += v;
#endif
}

// votes[l] += the sum of v over the warp: one shared-memory atomic per warp instead of one per lane.  Every lane of
// the warp must call it (lanes with nothing to add pass 0).
PL_HD void vote_add_warp(int* votes, int l, int v) {
#if PL_DEVICE_BUILD
  const int s = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0 && s != 0) atomicAdd(&votes[l], s);
#else
  votes[l] += v;
#endif
}

// raw weighted cross moment sum_i c_i x~_ip x~_iq from the tile store
PL_HD double gram_raw(const ModelView& M, const double* G, int p, int q) {
  int t = M.tile_of[(p >> 3) * M.ns + (q >> 3)];
  if (t >= 0) return G[(size_t)t * TILE + (p & 7) * SLOT + (q & 7)];
  t = -(t + 2);
  return G[(size_t)t * TILE + (q & 7) * SLOT + (p & 7)];
}

// In-place Cholesky of the lower triangle (row-major, leading dimension ld). false if not PD.
PL_HD bool chol_factor(double* A, int n, int ld) {
  for (int j = 0; j < n; ++j) {
    double diag0 = A[j * ld + j];
    double s = diag0;
    for (int k = 0; k < j; ++k) s -= A[j * ld + k] * A[j * ld + k];
    if (!(s > 1e-13 * fabs(diag0)) || !(diag0 > 0.0)) return false;
    double r = sqrt(s);
    A[j * ld + j] = r;
    for (int i = j + 1; i < n; ++i) {
      double t = A[i * ld + j];
      for (int k = 0; k < j; ++k) t -= A[i * ld + k] * A[j * ld + k];
      A[i * ld + j] = t / r;
    }
  }
  return true;
}

PL_HD void chol_solve(const double* A, int n, int ld, double* b) {
  for (int i = 0; i < n; ++i) {
    double t = b[i];
    for (int k = 0; k < i; ++k) t -= A[i * ld + k] * b[k];
    b[i] = t / A[i * ld + i];
  }
  for (int i = n - 1; i >= 0; --i) {
    double t = b[i];
    for (int k = i + 1; k < n; ++k) t -= A[k * ld + i] * b[k];
    b[i] = t / A[i * ld + i];
  }
}

// Minimum-norm solution of A x = b for a symmetric positive SEMI-definite A (full n x n storage) and b in its range
// -- the normal equations of a rank-deficient least-squares problem, where the reference's SVD-based solvers
// (scipy lstsq / gelsd in mode.py:50-52, statsmodels' pinv in scheme.py:50 and inner_model.py:76-77) return the
// minimum-norm coefficients.  Conjugate gradients started at 0 stay in range(A) and converge to exactly that solution
// in at most rank(A) steps; the tiny spurious eigenvalues of the null space are never resolved because the iteration
// stops on the residual.  t: 3 n doubles of scratch.  x must not alias b.
PL_HD void cg_min_norm(const double* A, int n, const double* b, double* x, double* t) {
  double *r = t, *p = t + n, *Ap = t + 2 * n;
  double rs = 0.0;
  for (int i = 0; i < n; ++i) { x[i] = 0.0; r[i] = b[i]; p[i] = b[i]; rs += b[i] * b[i]; }
  const double stop = 1e-28 * rs;
  for (int it = 0; it < 2 * n && rs > stop; ++it) {
    double pAp = 0.0;
    for (int i = 0; i < n; ++i) {
      double a = 0.0;
      for (int j = 0; j < n; ++j) a += A[i * n + j] * p[j];
      Ap[i] = a;
      pAp += p[i] * a;
    }
    if (!(pAp > 0.0)) break;
    const double alpha = rs / pAp;
    double rs_new = 0.0;
    for (int i = 0; i < n; ++i) {
      x[i] += alpha * p[i];
      r[i] -= alpha * Ap[i];
      rs_new += r[i] * r[i];
    }
    const double beta = rs_new / rs;
    for (int i = 0; i < n; ++i) p[i] = r[i] + beta * p[i];
    rs = rs_new;
  }
}

// Mode B (mode.py:50-52): the block matrix S_ll is factored once per replicate.  C: mode_b_scratch_doubles(k); `fill`
// writes element (r, c).  A rank-deficient block keeps the matrix itself and sets the flag C[k*k].
#define PL_MODE_B_PREPARE(C, k, ELEM)                                               \
  do {                                                                              \
    for (int r_ = 0; r_ < (k); ++r_)                                                \
      for (int c_ = 0; c_ <= r_; ++c_) (C)[r_ * (k) + c_] = ELEM(r_, c_);           \
    (C)[(k) * (k)] = 0.0;                                                           \
    if (!chol_factor((C), (k), (k))) {                                              \
      for (int r_ = 0; r_ < (k); ++r_)                                              \
        for (int c_ = 0; c_ < (k); ++c_) (C)[r_ * (k) + c_] = ELEM(r_, c_);         \
      (C)[(k) * (k)] = 1.0;                                                         \
    }                                                                               \
  } while (0)

// w <- S_ll^-1 w, or the minimum-norm least-squares weights when the block is rank deficient
PL_HD void mode_b_solve(double* C, int k, double* w) {
  if (C[k * k] == 0.0) {
    chol_solve(C, k, k, w);
    return;
  }
  double* x = C + k * k + 1;
  cg_min_norm(C, k, w, x, x + k);
  for (int i = 0; i < k; ++i) w[i] = x[i];
}

// Regression of LV i on its predecessors from a correlation matrix R (L x L): beta = R_pp^-1 R_pi, or the
// minimum-norm coefficients when R_pp is singular (collinear scores).  scratch: ols_scratch_doubles(max_deg), beta_out =
// scratch + max_deg^2.  Returns false only for non-finite input.
PL_HD bool regress_on_predecessors(const ModelView& M, const double* R, int i, double* scratch, double* beta_out) {
  int b0 = M.pred_begin[i], n = M.pred_begin[i + 1] - b0;
  double* A = scratch;
  for (int a = 0; a < n; ++a)
    for (int c = 0; c <= a; ++c) A[a * n + c] = R[M.pred_idx[b0 + a] * M.L + M.pred_idx[b0 + c]];
  for (int a = 0; a < n; ++a) beta_out[a] = R[M.pred_idx[b0 + a] * M.L + i];
  if (chol_factor(A, n, n)) {
    chol_solve(A, n, n, beta_out);
    return true;
  }
  double* rhs = beta_out + M.max_deg;  // four spare vectors follow beta in the scratch
  bool finite = true;
  for (int a = 0; a < n; ++a) {
    rhs[a] = beta_out[a];
    finite = finite && (rhs[a] - rhs[a] == 0.0);
    for (int c = 0; c < n; ++c) {
      A[a * n + c] = R[M.pred_idx[b0 + a] * M.L + M.pred_idx[b0 + c]];
      finite = finite && (A[a * n + c] - A[a * n + c] == 0.0);
    }
  }
  if (!finite) return false;
  cg_min_norm(A, n, rhs, beta_out, rhs + M.max_deg);
  return true;
}

// Shared-memory layout (doubles): see HostModel::solver_core_smem_doubles().
PL_HD void solve_replicate(const SolveArgs& A, double* smem) {
  const ModelView& M = A.M;
  const int L = M.L, Ppad = M.Ppad, tid = PL_TID, nt = PL_NT;
  double* w = smem;                  // [Ppad] current outer weights (un-normalised)
  double* u = w + Ppad;              // [Ppad] new weights / temporaries (the previous weights are w itself: the loop
                                     //        leaves w == w_old at its end, so no third array)
  double* m = u + Ppad;              // [Ppad] replicate column means of x~
  double* V = m + Ppad;              // [n_v]
  double* dinv = V + M.n_v;          // [L] 1/sd_pop(Y_l)
  double* sgn = dinv + L;            // [L]
  double* r2 = sgn + L;              // [L]
  double* bsum = r2 + L;             // [L]
  double* R = bsum + L;              // [L*L]  score correlations; after the inner model: total effects
  double* E = R + (size_t)L * L;     // [L*L]  inner weights (dead after the iteration)
  double* Bm = E;                    //        ... then the path coefficients
  double* red = E + (size_t)L * L;   // [40] reduction scratch + flags
  int* flag = (int*)(red + 34);      // [0] status
  int* votes = (int*)(red + 40);     // [L] ints
  int* unc = votes + 2 * ((L + 1) / 2);  // [L] ints: undecided voters (phase 3)

  const double N = A.N, invN = 1.0 / A.N;
  const double a_fac = (N - 1.0) / N;  // 1/correction^2, weights.py:44 (treat(..)/correction)

  for (int p = tid; p < Ppad; p += nt) m[p] = A.colsum[p] * invN;
  if (tid == 0) { flag[0] = STATUS_OK; flag[1] = 0; }
  PL_SYNC();

  double iss = 1.0;  // 1 / pooled scale^2
#define PL_S(p, q) ((gram_raw(M, A.G, (p), (q)) * invN - m[p] * m[q]) * iss)
  const int ols_stride = ols_scratch_doubles(M.max_deg);
  double* ols = A.ws + (M.ws_doubles - L * ols_stride);
  int iteration = 0, status = STATUS_OK;
  if (A.resume && A.state && A.phase != 1) {
    // the weights converged in phase 1 (same replicate, same moments): take w, V, dinv, R as they were left
    const double* S = A.state;
    for (int p = tid; p < Ppad; p += nt) w[p] = S[p];
    for (int v = tid; v < M.n_v; v += nt) V[v] = S[Ppad + v];
    for (int l = tid; l < L; l += nt) dinv[l] = S[Ppad + M.n_v + l];
    for (int e = tid; e < L * L; e += nt) R[e] = S[Ppad + M.n_v + L + e];
    const double* misc = S + Ppad + M.n_v + L + (size_t)L * L;
    iss = misc[0];
    iteration = (int)misc[1];
    status = (int)misc[2];
    PL_SYNC();
  } else {
  // pooled scale (config.py:302-303): sd over all N*P raw values, ddof=1, times sqrt((N-1)/N)
  if (M.scaled) {
    double ss = 0.0, gs = 0.0;
    for (int p = tid; p < Ppad; p += nt)
      if (M.col_lv[p] >= 0) {
        ss += gram_raw(M, A.G, p, p) - A.colsum[p] * A.colsum[p] * invN;
        gs += A.mu[p] + m[p];
      }
    ss = block_sum(ss, red);
    gs = block_sum(gs, red);
    double grand = gs / (double)M.P, dev = 0.0;
    for (int p = tid; p < Ppad; p += nt)
      if (M.col_lv[p] >= 0) {
        double d = A.mu[p] + m[p] - grand;
        dev += d * d;
      }
    dev = block_sum(dev, red);
    double var1 = (ss + N * dev) / (N * (double)M.P - 1.0);
    iss = 1.0 / (var1 * a_fac);
  }

  // ---- initial weights (weights.py:28-34): 1/sd of the block sum (correction cancels) ----------
  for (int l = tid; l < L; l += nt) {
    double bs = 0.0;
    int o = M.lv_off[l], k = M.lv_k[l];
    for (int r = 0; r < k; ++r)
      for (int c = 0; c < k; ++c) bs += PL_S(o + r, o + c);
    double wi = 1.0 / sqrt(bs);
    for (int r = 0; r < SLOT * ((k + SLOT - 1) / SLOT); ++r) {
      w[o + r] = (r < k) ? wi : 0.0;
      u[o + r] = 0.0;
    }
    // Mode B: factor S_ll once per replicate (the block Gram is iteration-invariant)
    if (M.lv_mode[l] == MODE_B) {
      double* C = A.ws + M.chol_b_off[l];
#define PL_BLK(r_, c_) PL_S(o + (r_), o + (c_))
      PL_MODE_B_PREPARE(C, k, PL_BLK);
#undef PL_BLK
    }
  }
  PL_SYNC();

  bool finalize = false;
  for (;;) {
    // ---- V_d = S_lj w_j for every needed directed LV pair  (Y = X W, weights.py:43) -----------
    for (int t = tid; t < M.n_pairs * M.kmax; t += nt) {
      int d = t / M.kmax, r = t - d * M.kmax;
      int l = M.pair_l[d], j = M.pair_j[d];
      if (r >= M.lv_k[l]) continue;
      int kj = M.lv_k[j], ol = M.lv_off[l], oj = M.lv_off[j];
      // sum_c S(p, oj + c) w_c with S = (G/N - m m') iss, slot by slot of block j (blocks start on slot boundaries):
      // one tile lookup per slot instead of one per element
      const int pp = ol + r;
      double g = 0.0, mw = 0.0;
      for (int c0 = 0; c0 < kj; c0 += SLOT) {
        const int tt = M.tile_of[(pp >> 3) * M.ns + ((oj + c0) >> 3)];
        const double* T = tt >= 0 ? A.G + (size_t)tt * TILE + (pp & 7) * SLOT : A.G + (size_t)(-(tt + 2)) * TILE + (pp & 7);
        const int step = tt >= 0 ? 1 : SLOT;
        const int cn = kj - c0 < SLOT ? kj - c0 : SLOT;
        for (int c = 0; c < cn; ++c) {
          const double wc = w[oj + c0 + c];
          g += T[c * step] * wc;
          mw += m[oj + c0 + c] * wc;
        }
      }
      V[M.pair_voff[d] + r] = (g * invN - m[pp] * mw) * iss;
    }
    PL_SYNC();
    // ---- population variance of Y_l, standardisation factor (weights.py:44) ---------------------
    for (int l = tid; l < L; l += nt) {
      int d = M.lv_pair_begin[l], o = M.lv_off[l];
      double var = 0.0;
      for (int r = 0; r < M.lv_k[l]; ++r) var += w[o + r] * V[M.pair_voff[d] + r];
      dinv[l] = 1.0 / sqrt(var);
      R[l * L + l] = 1.0;
    }
    PL_SYNC();
    // ---- correlations of the LV scores for the needed pairs -------------------------------------
    for (int d = tid; d < M.n_pairs; d += nt) {
      int l = M.pair_l[d], j = M.pair_j[d];
      if (l > j) {
        int o = M.lv_off[l];
        double acc = 0.0;
        for (int r = 0; r < M.lv_k[l]; ++r) acc += w[o + r] * V[M.pair_voff[d] + r];
        acc *= dinv[l] * dinv[j];
        R[l * L + j] = acc;
        R[j * L + l] = acc;
      }
    }
    PL_SYNC();
    if (finalize) break;

    // ---- inner weights E (scheme.py:27-28, 36-37, 45-54); E[j*L+l] multiplies Y_j in Z_l --------
    if (A.scheme == SCHEME_PATH) {
      for (int e = tid; e < L * L; e += nt) E[e] = 0.0;
      PL_SYNC();
      for (int i = tid; i < L; i += nt) {
        int n = M.pred_begin[i + 1] - M.pred_begin[i];
        if (n > 0) {
          double* sc = ols + (size_t)i * ols_stride;
          double* beta = sc + M.max_deg * M.max_deg;
          if (!regress_on_predecessors(M, R, i, sc, beta)) flag[0] = STATUS_SINGULAR;
          for (int a = 0; a < n; ++a) E[M.pred_idx[M.pred_begin[i] + a] * L + i] = beta[a];
        }
        for (int a = M.succ_begin[i]; a < M.succ_begin[i + 1]; ++a) {
          int k = M.succ_idx[a];
          E[k * L + i] = R[k * L + i];
        }
      }
    } else {
      for (int e = tid; e < L * L; e += nt) {
        int j = e / L, l = e - j * L;
        double val = 0.0;
        if (M.path[j * L + l] | M.path[l * L + j]) {
          double r = R[e];
          if (A.scheme == SCHEME_CENTROID) val = (r > 0.0) ? 1.0 : ((r < 0.0) ? -1.0 : r);  // numpy.sign
          else val = a_fac * r;  // cov ddof=1 of scores scaled by a (quirk Q3)
        }
        E[e] = val;
      }
    }
    PL_SYNC();
    // ---- outer weights (weights.py:46-50): (1/N) X_l' Z_l, Z_l = sum_j a d_j E[j,l] Y_j ----------
    for (int p = tid; p < Ppad; p += nt) {
      int l = M.col_lv[p];
      if (l < 0) continue;
      int r = p - M.lv_off[l];
      double acc = 0.0;
      for (int d = M.lv_pair_begin[l] + 1; d < M.lv_pair_begin[l + 1]; ++d) {
        int j = M.pair_j[d];
        double e = E[j * L + l];
        if (e != 0.0) acc += a_fac * dinv[j] * e * V[M.pair_voff[d] + r];
      }
      u[p] = acc;
    }
    PL_SYNC();
    for (int l = tid; l < L; l += nt)
      if (M.lv_mode[l] == MODE_B && flag[0] == STATUS_OK)
        mode_b_solve(A.ws + M.chol_b_off[l], M.lv_k[l], u + M.lv_off[l]);  // mode.py:50-52
    PL_SYNC();
    // ---- convergence (weights.py:51-53) ----------------------------------------------------------
    double part = 0.0;
    for (int p = tid; p < Ppad; p += nt) {
      double df = fabs(w[p]) - fabs(u[p]);
      part += df * df;
      w[p] = u[p];
    }
    double conv = block_sum(part, red);
    ++iteration;
    PL_SYNC();
    // weights.py:183  (a NaN criterion keeps iterating until the cap, like the reference)
    if ((conv < A.tol) || (iteration > A.max_iter) || flag[0] != STATUS_OK) finalize = true;
  }
  status = flag[0];
  if (status == STATUS_OK && iteration > A.max_iter) status = STATUS_NOT_CONVERGED;  // weights.py:185 (Q4)
  }  // (not resumed)

  // ---- final normalisation, sign vote (weights.py:56-68) --------------------------------------------
  // after the loop V, dinv and R hold the values for the final weights
  for (int p = tid; p < Ppad; p += nt) {
    int l = M.col_lv[p];
    u[p] = (l >= 0) ? w[p] * dinv[l] : 0.0;  // wf: scores have unit population variance
  }
  for (int l = tid; l < L; l += nt) votes[l] = 0;
  PL_SYNC();  // u is read across threads below (racecheck: shared-memory RAW hazard without this)
  if (A.phase == 1) {  // sparse tile set, first pass: hand the weights to the cross-moment kernel
    for (int p = tid; p < Ppad; p += nt) A.wf_out[p] = u[p];
    if (A.sh_out)
      for (int l = tid; l < L; l += nt) {
        double sh = 0.0;
        for (int r = 0; r < M.lv_k[l]; ++r) sh += m[M.lv_off[l] + r] * u[M.lv_off[l] + r];
        A.sh_out[l] = sh;
      }
    if (A.state) {
      double* S = A.state;
      for (int p = tid; p < Ppad; p += nt) S[p] = w[p];
      for (int v = tid; v < M.n_v; v += nt) S[Ppad + v] = V[v];
      for (int l = tid; l < L; l += nt) S[Ppad + M.n_v + l] = dinv[l];
      for (int e = tid; e < L * L; e += nt) S[Ppad + M.n_v + L + e] = R[e];
      if (tid == 0) {
        double* misc = S + Ppad + M.n_v + L + (size_t)L * L;
        misc[0] = iss;
        misc[1] = (double)iteration;
        misc[2] = (double)status;
      }
    }
    if (tid == 0) {
      if (A.iters) *A.iters = iteration;
      if (A.status) *A.status = status;
    }
    return;
  }
  if (A.phase == 2)  // mean of the un-centred score: sum_q m_q wf_q (bsum is free after the init)
    for (int l = tid; l < L; l += nt) {
      double sh = 0.0;
      for (int r = 0; r < M.lv_k[l]; ++r) sh += m[M.lv_off[l] + r] * u[M.lv_off[l] + r];
      bsum[l] = sh;
    }
  if (A.phase == 3) {
    // per-column factor of the vote's error bound (w is dead once u = w dinv is formed)
    for (int p = tid; p < Ppad; p += nt) w[p] = M.col_lv[p] >= 0 ? sqrt(gram_raw(M, A.G, p, p)) * A.inv_sd[p] : 0.0;
    for (int l = tid; l < L; l += nt) {
      unc[l] = 0;
      // fp32 score generation: |t^ - t| <= (k+4) 2^-24 (sum_k |x_k w_k| + |sh|) per row, hence (Cauchy-Schwarz
      // over the rows)  sum_i c_i xh_ip |t^ - t| <= sqrt(sum c xh_p^2) * bsum[l] with
      // bsum[l] = (k+4) 2^-24 sqrt(2 (|w_l|^2 sum_q G_qq + N sh_l^2))
      double w2 = 0.0, g = 0.0, sh = 0.0;
      for (int r = 0; r < M.lv_k[l]; ++r) {
        const int q = M.lv_off[l] + r;
        w2 += u[q] * u[q];
        g += gram_raw(M, A.G, q, q);
        sh += m[q] * u[q];
      }
      bsum[l] = (M.lv_k[l] + 4) * 6.0e-8 * sqrt(2.0 * (w2 * g + N * sh * sh));
      if (A.fast_uncentred) {
        // fused kernel: t'^ = fp32 sum over the block of fp16(xh_q) * fp16(w'_q), xh = x~/sd, w' = wf sd:
        // |t'^ - t'| <= 1.0e-3 |xh_blk| |w'_blk| per row (two fp16 roundings 2^-11 each, one truncating fp32
        // accumulation step per 16 columns), hence over the rows (Cauchy-Schwarz)
        //   sum_i c_i |xh_ip| |t'^ - t'| <= sqrt(sum c xh_p^2) * 1.0e-3 |w'_l| sqrt(sum_q G_qq / sd_q^2)
        double wp2 = 0.0, gh = 0.0;
        for (int r = 0; r < M.lv_k[l]; ++r) {
          const int q = M.lv_off[l] + r;
          const double isd = A.inv_sd[q];
          if (isd > 0.0) {
            wp2 += u[q] * u[q] / (isd * isd);
            gh += gram_raw(M, A.G, q, q) * isd * isd;
          }
        }
        bsum[l] = 1.0e-3 * sqrt(wp2 * gh);
        sgn[l] = sh;  // (sgn is free until the votes are counted) mean of the un-centred score
      }
      // the whole LV-side factor of the bound: E = sum_i xh_ip (c_i t_il) has the sign of cov(x_p, score_l); fp16
      // operands + truncating fp32 tensor-core accumulation give |fl(E) - E| <= gamma sqrt(sum c xh^2) sqrt(sum c t^2)
      // (Cauchy-Schwarz), sum c t^2 = N / iss (unit-variance scores on the treated scale) + N sh^2 when the scores
      // were not centred.  2e-3 covers the fp16 roundings of both MMA operands (2^-10), <= 1100 truncating fp32
      // accumulation steps per CTA (2^-22 each, measured 1e-7) and the fp32 adds that join the row ranges.
      bsum[l] = 2.0e-3 * sqrt(N / iss + (A.fast_uncentred ? N * sh * sh : 0.0)) + bsum[l];
    }
  }
  PL_SYNC();
  if (A.phase == 3) {
    // thread = manifest variable, loop over the LVs: the per-column factors stay in registers, the strided loads of
    // the cross moments are independent of each other, votes are summed over the warp before they touch shared memory
    for (int p0 = 0; p0 < Ppad; p0 += nt) {
      const int p = p0 + tid;
      const int lp = p < Ppad ? M.col_lv[p] : -1;
      const double colf = lp >= 0 ? w[p] : 0.0;
      const double shift = (lp >= 0 && A.fast_uncentred) ? A.colsum[p] * A.inv_sd[p] : 0.0;
      const float* cf = A.fast_cross + ((size_t)(lp >= 0 ? p : 0) * L) * A.fast_nb + A.fast_b;
      for (int l = 0; l < L; ++l) {
        int dv = 0, du = 0;
        if (lp >= 0) {
          if (!M.omega[lp * L + l]) {
            // E = E' - sh_l sum_i c_i xh_ip (exact in fp64) when the scores behind E' were not centred
            const double v = (double)cf[(size_t)l * A.fast_nb] - (A.fast_uncentred ? sgn[l] * shift : 0.0);
            if (v - v == 0.0 && fabs(v) > bsum[l] * colf) dv = v < 0.0 ? -1 : 1;
            else du = 1;
          } else {  // LV pair inside the tile set: exact covariance from the Gram tiles
            double acc = 0.0;
            const int o = M.lv_off[l];
            for (int c = 0; c < M.lv_k[l]; ++c) acc += PL_S(p, o + c) * u[o + c];
            if (A.crossloadings) A.crossloadings[(size_t)M.col_src[p] * L + l] = acc / sqrt(PL_S(p, p));
            if (acc == acc) dv = signbit(acc) ? -1 : 1;
          }
        }
        vote_add_warp(votes, l, dv);
        vote_add_warp(unc, l, du);
      }
    }
  } else
  // cov(x_p, score_l) for ALL (p, l): every manifest variable votes on every LV (quirk Q6)
  for (int t = tid; t < Ppad * L; t += nt) {
    int p = t / L, l = t - p * L;
    if (M.col_lv[p] < 0) continue;
    double acc = 0.0;
    if (A.phase == 2) {
      const double raw = A.cross[((size_t)(p >> 3) * M.ng + (l >> 3)) * TILE + (p & 7) * SLOT + (l & 7)];
      acc = (raw * invN - m[p] * bsum[l]) * iss;
    } else {
      int o = M.lv_off[l];
      for (int c = 0; c < M.lv_k[l]; ++c) acc += PL_S(p, o + c) * u[o + c];
    }
    if (A.crossloadings) A.crossloadings[(size_t)M.col_src[p] * L + l] = acc / sqrt(PL_S(p, p));
    // copysign(1, cor): +1 for cor >= +0, -1 for cor < 0 or -0 (weights.py:63-64); NaN does not vote
    if (acc == acc) vote_add(votes, l, signbit(acc) ? -1 : 1);
  }
  PL_SYNC();
  if (A.phase == 3) {
    // vote = votes + (something in [-unc, +unc]); decided when the interval does not straddle zero
    for (int l = tid; l < L; l += nt)
      if (!(votes[l] - unc[l] >= 0 || votes[l] + unc[l] < 0)) flag[1] = 1;
    PL_SYNC();
    if (flag[1]) {
      if (tid == 0) {
        if (A.iters) *A.iters = iteration;
        if (A.status) *A.status = (status == STATUS_OK) ? STATUS_AMBIGUOUS : status;
      }
      return;
    }
  }
  for (int l = tid; l < L; l += nt) sgn[l] = (votes[l] < 0) ? -1.0 : 1.0;
  PL_SYNC();
  // flip score correlations
  for (int e = tid; e < L * L; e += nt) {
    int i = e / L, j = e - i * L;
    R[e] *= sgn[i] * sgn[j];
    Bm[e] = 0.0;
  }
  PL_SYNC();
  // ---- inner model (inner_model.py:72-83): OLS with intercept on centred unit-variance scores ----
  for (int i = tid; i < L; i += nt) {
    int n = M.pred_begin[i + 1] - M.pred_begin[i];
    double rr = 0.0;
    if (n > 0) {
      double* sc = ols + (size_t)i * ols_stride;
      double* beta = sc + M.max_deg * M.max_deg;
      if (!regress_on_predecessors(M, R, i, sc, beta)) {
        if (status == STATUS_OK) flag[0] = STATUS_SINGULAR;
      } else {
        for (int a = 0; a < n; ++a) {
          int j = M.pred_idx[M.pred_begin[i] + a];
          Bm[i * L + j] = beta[a];
          rr += beta[a] * R[j * L + i];
        }
      }
    }
    r2[i] = rr;
  }
  PL_SYNC();
  if (status == STATUS_OK) status = flag[0];
  // ---- total effects (inner_model.py:33-49): T = B + B T, column by column ----------------------
  double* T = R;  // (the score correlations are dead: the regressions above were their last readers)
  for (int j = tid; j < L; j += nt)
    for (int i = 0; i < L; ++i) {
      double acc = Bm[i * L + j];
      for (int k = j + 1; k < i; ++k) acc += Bm[i * L + k] * T[k * L + j];
      T[i * L + j] = (i > j) ? acc : 0.0;
    }
  PL_SYNC();

  // ---- outputs -------------------------------------------------------------------------------------
  const int P = M.P, ne = M.n_eff;
  const double inv_scale = sqrt(iss);
  for (int p = tid; p < Ppad; p += nt) {
    int l = M.col_lv[p];
    if (l < 0) {
      if (A.score_coef) A.score_coef[p] = 0.0;
      continue;
    }
    int src = M.col_src[p];
    double wfp = u[p];
    // in-block loading = corr(x_p, flipped score_l) = sgn * (S_ll wf)_p / sd(x_p)   (bootstrap.py:65-66)
    double load = sgn[l] * V[M.pair_voff[M.lv_pair_begin[l]] + (p - M.lv_off[l])] * dinv[l] / sqrt(PL_S(p, p));
    if (A.out_row) {
      A.out_row[src] = wfp;
      A.out_row[P + L + 2 * ne + src] = load;
    }
    if (A.weights) A.weights[src] = wfp;
    if (A.loadings) A.loadings[src] = load;
    if (A.score_coef) A.score_coef[p] = sgn[l] * wfp * inv_scale;
  }
  for (int l = tid; l < L; l += nt) {
    if (A.out_row) A.out_row[P + l] = r2[l];
    if (A.r2) A.r2[l] = r2[l];
    if (A.score_shift) {
      double sh = 0.0;
      for (int r = 0; r < M.lv_k[l]; ++r) sh += m[M.lv_off[l] + r] * sgn[l] * u[M.lv_off[l] + r] * inv_scale;
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
  if (A.crossloadings) {
    PL_SYNC();
    for (int t = tid; t < P * L; t += nt) A.crossloadings[t] *= sgn[t % L];
  }
  if (tid == 0) {
    if (A.iters) *A.iters = iteration;
    if (A.status) *A.status = status;
  }
#undef PL_S
}

}  // namespace plspm
