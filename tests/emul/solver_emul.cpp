// TEST INFRASTRUCTURE ONLY -- never loaded by the product package.
//
// Host emulation of the device pipeline (slot layout + global centring, resample counts,
// weighted Gram tiles, per-replicate solver, scores) so that tests/ can check the solver
// logic in csrc/solver_core.h against the oracle in a container without a GPU.  The Gram
// tiles are produced here by a plain triple loop that writes the same tile layout the
// sm_100a Gram kernel writes; the solver is the shipped source compiled for the host.
#define PLSPM_HOST_EMUL 1
#include <cstdint>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../plspm-python_b200/csrc/plspm_model.h"
#include "../../plspm-python_b200/csrc/solver_core.h"
#include "../../plspm-python_b200/csrc/solver_num.h"

using namespace plspm;

extern "C" int emul_model_info(int L, const int32_t* block_sizes, const int8_t* modes, const int8_t* path, int scaled,
                               int tile_policy, int32_t* info /*[8]*/, int32_t* eff_from, int32_t* eff_to) {
  HostModel m;
  std::string err;
  if (build_model(L, block_sizes, modes, path, scaled, tile_policy, m, err)) return 1;
  info[0] = m.P; info[1] = m.Ppad; info[2] = m.n_tiles; info[3] = m.n_tg; info[4] = m.n_pairs;
  info[5] = m.n_eff; info[6] = m.n_out(); info[7] = m.full;
  if (eff_from)
    for (int e = 0; e < m.n_eff; ++e) { eff_from[e] = m.eff_from[e]; eff_to[e] = m.eff_to[e]; }
  return 0;
}

static int g_vote_mode = 0;      // 0: exact cross moments; 1: emulate the low-precision vote (phase 3)
static int g_last_ambiguous = 0;
static int g_resume = 1;         // 1: phases 2 / 3 resume from the state phase 1 saved (what the CUDA library does); 0: iterate again
extern "C" void emul_set_vote_mode(int mode) { g_vote_mode = mode; }
extern "C" void emul_set_resume(int on) { g_resume = on; }
extern "C" int emul_last_ambiguous() { return g_last_ambiguous; }

extern "C" int emul_fit(int L, const int32_t* block_sizes, const int8_t* modes, const int8_t* path, int scaled,
                        int tile_policy, const double* X, int64_t N, const int32_t* idx, int scheme, double tol,
                        int max_iter, double* out_row, double* weights, double* loadings, double* r2, double* paths,
                        double* total, double* crossloadings, double* scores, int32_t* iters, int32_t* status) {
  HostModel m;
  std::string err;
  if (build_model(L, block_sizes, modes, path, scaled, tile_policy, m, err)) return 1;
  const int P = m.P, Pp = m.Ppad;
  // upload-time layout: slot padded, globally centred
  std::vector<double> mu(Pp, 0.0), Xs((size_t)N * Pp, 0.0);
  for (int p = 0; p < P; ++p) {
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) s += X[i * P + p];
    mu[m.src_col[p]] = s / (double)N;
  }
  for (int64_t i = 0; i < N; ++i)
    for (int p = 0; p < P; ++p) Xs[i * Pp + m.src_col[p]] = X[i * P + p] - mu[m.src_col[p]];
  std::vector<double> cnt(N, idx ? 0.0 : 1.0);
  if (idx)
    for (int64_t i = 0; i < N; ++i) cnt[idx[i]] += 1.0;
  std::vector<double> G((size_t)m.n_tiles * TILE, 0.0), colsum(Pp, 0.0);
  for (int64_t i = 0; i < N; ++i) {
    if (cnt[i] == 0.0) continue;
    const double* x = &Xs[i * Pp];
    for (int p = 0; p < Pp; ++p) colsum[p] += cnt[i] * x[p];
    for (int t = 0; t < m.n_tiles; ++t) {
      const double* xa = x + m.tile_sa[t] * SLOT;
      const double* xb = x + m.tile_sb[t] * SLOT;
      double* g = &G[(size_t)t * TILE];
      for (int r = 0; r < SLOT; ++r)
        for (int c = 0; c < SLOT; ++c) g[r * SLOT + c] += xa[r] * (cnt[i] * xb[c]);
    }
  }
  std::vector<double> smem(m.solver_core_smem_doubles(), 0.0), ws(m.ws_doubles, 0.0), coef(Pp, 0.0), shift(L, 0.0);
  SolveArgs A;
  std::memset(&A, 0, sizeof(A));
  A.M = m.host_view();
  A.G = G.data(); A.colsum = colsum.data(); A.mu = mu.data(); A.N = (double)N;
  A.scheme = scheme; A.tol = tol; A.max_iter = max_iter; A.ws = ws.data();
  A.out_row = out_row; A.weights = weights; A.loadings = loadings; A.r2 = r2; A.paths = paths; A.total = total;
  A.crossloadings = crossloadings; A.score_coef = coef.data(); A.score_shift = shift.data();
  A.iters = iters; A.status = status;
  std::vector<double> wf(Pp, 0.0), cross((size_t)m.n_cross * TILE, 0.0);
  if (m.full) {
    A.phase = 0;
    solve_replicate(A, smem.data());
  } else {
    // sparse tile set: weights first, then the P x L cross-moment pass, then the full solve
    std::vector<double> sh(L, 0.0), state(solver_state_doubles(A.M), 0.0);
    A.phase = 1; A.wf_out = wf.data(); A.sh_out = sh.data();
    if (g_resume) { A.state = state.data(); A.resume = 1; }
    solve_replicate(A, smem.data());
    g_last_ambiguous = 0;
    bool need_exact = true;
    if (g_vote_mode == 1) {
      // emulate the fp16 tensor-core pass: E[p][l] = sum_i xh_ip c_i t_il with a perturbation of half the
      // guaranteed error bound (alternating sign), stored as float
      std::vector<double> inv_sd(Pp, 0.0), sumx2(Pp, 0.0), sumt2(L, 0.0);
      for (int p = 0; p < Pp; ++p) {
        double q = 0.0;
        for (int64_t i = 0; i < N; ++i) q += Xs[i * Pp + p] * Xs[i * Pp + p];
        inv_sd[p] = q > 0 ? 1.0 / std::sqrt(q / (double)N) : 0.0;
      }
      std::vector<double> E((size_t)Pp * L, 0.0);
      for (int64_t i = 0; i < N; ++i) {
        if (cnt[i] == 0.0) continue;
        const double* x = &Xs[i * Pp];
        for (int p = 0; p < Pp; ++p) sumx2[p] += cnt[i] * x[p] * inv_sd[p] * x[p] * inv_sd[p];
        for (int l = 0; l < L; ++l) {
          double t = -sh[l];
          for (int c = m.lv_off[l]; c < m.lv_off[l + 1]; ++c) t += x[c] * wf[c];
          sumt2[l] += cnt[i] * t * t;
          for (int p = 0; p < Pp; ++p) E[(size_t)p * L + l] += x[p] * inv_sd[p] * cnt[i] * t;
        }
      }
      std::vector<float> Ef((size_t)Pp * L);
      for (int p = 0; p < Pp; ++p)
        for (int l = 0; l < L; ++l) {
          double pert = 1.0e-3 * std::sqrt(sumx2[p]) * std::sqrt(sumt2[l]) * (((p + l) & 1) ? 1.0 : -1.0);
          Ef[(size_t)p * L + l] = (float)(E[(size_t)p * L + l] + pert);
        }
      std::fill(smem.begin(), smem.end(), 0.0);
      A.phase = 3; A.fast_cross = Ef.data(); A.inv_sd = inv_sd.data(); A.fast_nb = 1; A.fast_b = 0;
      int st_fast = 0;
      int32_t* user_status = A.status;
      A.status = &st_fast;
      solve_replicate(A, smem.data());
      A.status = user_status;
      if (st_fast == STATUS_AMBIGUOUS) g_last_ambiguous = 1;
      else { need_exact = false; if (status) *status = st_fast; }
    }
    if (need_exact)
    for (int64_t i = 0; i < N; ++i) {
      if (cnt[i] == 0.0) continue;
      const double* x = &Xs[i * Pp];
      for (int l = 0; l < L; ++l) {
        double sc = 0.0;
        for (int c = m.lv_off[l]; c < m.lv_off[l + 1]; ++c) sc += x[c] * wf[c];
        sc *= cnt[i];
        for (int p = 0; p < Pp; ++p)
          cross[((size_t)(p >> 3) * m.ng + (l >> 3)) * TILE + (p & 7) * SLOT + (l & 7)] += x[p] * sc;
      }
    }
    if (need_exact) {
      std::fill(smem.begin(), smem.end(), 0.0);
      A.phase = 2; A.cross = cross.data();
      solve_replicate(A, smem.data());
    }
  }
  if (scores && !idx)
    for (int64_t i = 0; i < N; ++i)
      for (int l = 0; l < L; ++l) {
        double s = 0.0;
        for (int c = m.lv_off[l]; c < m.lv_off[l] + m.lv_k[l]; ++c) s += Xs[i * Pp + c] * coef[c];
        scores[i * L + l] = s - shift[l];
      }
  return 0;
}


// Non-metric (Scale.NUM / RAW) fit: host-driven outer loop around num_step(), with the score criterion
// sum_{i,l} c_i (|y_old| - |y_new|)^2 evaluated per observation on the host (the CUDA library does this in
// conv_kernel).
static int g_num_decomposition_errors = 0;
extern "C" int emul_num_decomposition_errors() { return g_num_decomposition_errors; }
extern "C" int emul_fit_num(int L, const int32_t* block_sizes, const int8_t* modes, const int8_t* path,
                            int tile_policy, const double* X, int64_t N, const int32_t* idx, int scheme, double tol,
                            int max_iter, double* out_row, double* weights, double* loadings, double* r2, double* paths,
                            double* total, double* crossloadings, double* scores, int32_t* iters, int32_t* status) {
  HostModel m;
  std::string err;
  if (build_model(L, block_sizes, modes, path, 0, tile_policy, m, err)) return 1;
  const int P = m.P, Pp = m.Ppad;
  std::vector<double> mu(Pp, 0.0), Xs((size_t)N * Pp, 0.0);
  for (int p = 0; p < P; ++p) {
    double s = 0.0;
    for (int64_t i = 0; i < N; ++i) s += X[i * P + p];
    mu[m.src_col[p]] = s / (double)N;
  }
  for (int64_t i = 0; i < N; ++i)
    for (int p = 0; p < P; ++p) Xs[i * Pp + m.src_col[p]] = X[i * P + p] - mu[m.src_col[p]];
  std::vector<double> cnt(N, idx ? 0.0 : 1.0);
  if (idx)
    for (int64_t i = 0; i < N; ++i) cnt[idx[i]] += 1.0;
  std::vector<double> G((size_t)m.n_tiles * TILE, 0.0), colsum(Pp, 0.0);
  for (int64_t i = 0; i < N; ++i) {
    if (cnt[i] == 0.0) continue;
    const double* x = &Xs[i * Pp];
    for (int p = 0; p < Pp; ++p) colsum[p] += cnt[i] * x[p];
    for (int t = 0; t < m.n_tiles; ++t) {
      const double* xa = x + m.tile_sa[t] * SLOT;
      const double* xb = x + m.tile_sb[t] * SLOT;
      double* g = &G[(size_t)t * TILE];
      for (int r = 0; r < SLOT; ++r)
        for (int c = 0; c < SLOT; ++c) g[r * SLOT + c] += xa[r] * (cnt[i] * xb[c]);
    }
  }
  std::vector<double> smem(m.solver_smem_doubles(), 0.0), ws(m.ws_doubles, 0.0), a(Pp, 0.0), co(Pp), cn(Pp), so(L), sn(L),
      coef(Pp, 0.0), shift(L, 0.0);
  int meta[4] = {0, 0, 0, 0};
  NumStepArgs A;
  std::memset(&A, 0, sizeof(A));
  A.M = m.host_view();
  A.G = G.data(); A.colsum = colsum.data(); A.N = (double)N; A.scheme = scheme; A.tol = tol; A.max_iter = max_iter;
  A.ws = ws.data(); A.a = a.data(); A.meta = meta;
  A.coef_old = co.data(); A.coef_new = cn.data(); A.shift_old = so.data(); A.shift_new = sn.data();
  A.out_row = out_row; A.weights = weights; A.loadings = loadings; A.r2 = r2; A.paths = paths; A.total = total;
  A.crossloadings = crossloadings; A.score_coef = coef.data(); A.score_shift = shift.data();
  A.iters = iters; A.status = status;
  A.conv_in = 0.0;
  double conv_main = 0.0;
  A.conv_main = &conv_main;
  for (int guard = 0; guard < max_iter + 5 && !meta[1]; ++guard) {
    num_step(A, smem.data());
    if (meta[1]) break;
    // criterion = second-moment part (from num_step) + 4 sum c y_old y_new over the rows whose score changes
    // sign; the direct per-observation evaluation is kept as a cross-check of the decomposition
    double conv = conv_main, direct = 0.0;
    for (int64_t i = 0; i < N; ++i) {
      if (cnt[i] == 0.0) continue;
      const double* x = &Xs[i * Pp];
      for (int l = 0; l < L; ++l) {
        double yo = -so[l], yn = -sn[l];
        for (int c = m.lv_off[l]; c < m.lv_off[l + 1]; ++c) { yo += x[c] * co[c]; yn += x[c] * cn[c]; }
        if (yo * yn < 0.0) conv += 4.0 * cnt[i] * yo * yn;
        const double df = std::fabs(yo) - std::fabs(yn);
        direct += cnt[i] * df * df;
      }
    }
    if (std::fabs(conv - direct) > 1e-9 * (std::fabs(direct) + 1e-12) + 1e-18) g_num_decomposition_errors++;
    A.conv_in = conv;
  }
  if (scores && !idx)
    for (int64_t i = 0; i < N; ++i)
      for (int l = 0; l < L; ++l) {
        double s = 0.0;
        for (int c = m.lv_off[l]; c < m.lv_off[l] + m.lv_k[l]; ++c) s += Xs[i * Pp + c] * coef[c];
        scores[i * L + l] = s - shift[l];
      }
  return 0;
}
