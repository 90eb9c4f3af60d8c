// Metric path: solve_kernel (solver_core.h, one CTA per replicate) and scores_kernel.
// Part of the single translation unit plspm_b200.cu (included there, in this order); see DESIGN.md §4.
#pragma once

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
  double* state; int64_t state_stride; int resume;  // phase 1 -> phases 2 / 3 (see SolveArgs::state)
  int phase;                             // see SolveArgs::phase
  double* wf;                            // [nrep][Ppad] (sparse tile sets)
  const double* cross; int64_t cross_stride;
  const float* fast_cross; const double* inv_sd; int64_t fast_nb; int fast_uncentred;  // phase 3
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
  A.fast_uncentred = b.fast_uncentred;
  A.sh_out = b.sh ? b.sh + rep * b.M.L : nullptr;
  A.ws = b.ws + rep * (int64_t)b.M.ws_doubles;
  A.state = b.state ? b.state + rep * b.state_stride : nullptr;
  A.resume = b.resume;
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
