// Non-metric path: one outer iteration per launch (solver_num.h) for every unfinished replicate.
// Part of the single translation unit plspm_b200.cu (included there, in this order); see DESIGN.md §4.
#pragma once

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
