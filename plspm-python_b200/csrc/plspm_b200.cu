// plspm_b200: sm_100a kernels + C ABI (include/plspm_b200.h).  This file holds the handles, the launch plans,
// the batch drivers and the C ABI; the kernels are in kernels_*.cuh (included below, one translation unit).
//
// Data flow of one bootstrap batch (replicates are the batch dimension):
//
//   counts_kernel     Philox4x32-10 (or injected) resample indices -> multiplicity c[b][i] (u32)
//   second moments    N >= 4096: counts8_kernel + ONE int8 GEMM on the tensor cores over digit planes of the
//                     pair products x~_ip x~_iq (exact int32 sums) + zcombine_kernel; same for the column sums
//                     (kernels_digits.cuh).  Otherwise gram_kernel / colsum_kernel in fp64 (kernels_gram.cuh):
//                     8x8 register tiles, X~ row tiles staged by the TMA engine into an mbarrier ring shared by
//                     the 8 (replicate, tile group) warps of a CTA.
//   solve_kernel      one CTA per replicate: the whole PLS-PM iteration in the covariance domain
//                     (solver_core.h), inner model, effects, loadings
//   sign vote         sparse tile sets: scoregen_kernel (fp32 scores -> fp16) + cuBLAS fp16 GEMM with a rigorous
//                     error bound, or the exact fp64 cross-moment pass (gram_kernel<true>)
//   non-metric path   num_step_kernel (solver_num.h) + conv_kernel per outer iteration, driven from the host
//   scores_kernel     single fit only: scores = X~ . coef - shift   (N x L, HBM-bound)
//
// Nothing here falls back to a CPU path: without a CUDA device every entry point fails.
#include <cublas_v2.h>
#include <cuda.h>
#include <type_traits>
#include <chrono>
#include <map>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/plspm_b200.h"
#include "plspm_model.h"
#include "solver_core.h"
#include "solver_num.h"

using namespace plspm;
#include "device_runtime.cuh"

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
struct plspm_model {
  HostModel h;
  int device = 0;
  bool numeric = false;  // non-metric treatment with numeric scales (plspm_model_set_numeric)
  std::vector<void*> dev_allocs;
  ModelView dv;  // device pointers
};

// cuBLAS handles, recycled per device (a data handle owns one exclusively while it lives)
static std::mutex g_blas_mu;
static std::map<int, std::vector<cublasHandle_t>> g_blas_free;
static cublasHandle_t blas_acquire(int device) {
  {
    std::lock_guard<std::mutex> lk(g_blas_mu);
    auto& v = g_blas_free[device];
    if (!v.empty()) {
      cublasHandle_t h = v.back();
      v.pop_back();
      return h;
    }
  }
  cublasHandle_t h = nullptr;
  return cublasCreate(&h) == CUBLAS_STATUS_SUCCESS ? h : nullptr;
}
static void blas_release(int device, cublasHandle_t h) {
  std::lock_guard<std::mutex> lk(g_blas_mu);
  g_blas_free[device].push_back(h);
}

struct Workspace {
  void* ptr = nullptr;
  size_t bytes = 0;
};

// Bootstrap with missing values (kernels_impute.cuh): the handle holds the AUGMENTED matrix [x0 | m] under an augmented
// model; replicates are solved under the base model on the moments of the imputed data.
struct ImputeCtx {
  const plspm_model* base = nullptr;
  int *ax = nullptr, *am = nullptr;  // [base Ppad] augmented padded column of x0_p and of its missing indicator (-1: none)
  double* mu_base = nullptr;         // [base Ppad] upload mean of the zero-filled column (the base moments are taken about it)
  double *G = nullptr, *colsum = nullptr, *ws = nullptr;  // moments of the imputed replicates + solver scratch, base layout
  int64_t cap = 0;                   // replicates the three buffers hold
};

struct plspm_data {
  const plspm_model* model = nullptr;
  int64_t N = 0;
  double* X = nullptr;   // [N][Ppad] slot layout, globally centred
  double* mu = nullptr;  // [Ppad]
  double* colsum0 = nullptr;  // [Ppad] column sums of the centred matrix (rounding residue of the centring): a single fit's column sums
  // low-precision copy for the tensor-core sign vote (sparse tile sets): xh = x~ / sd (fp16)
  __half* Xh = nullptr;      // [N][Ppad]
  float* Xf = nullptr;       // [N][Ppad] fp32 copy of x~ (score generation of the sign vote)
  double* inv_sd = nullptr;  // [Ppad]
  cublasHandle_t blas = nullptr;
  int blas_device = 0;
  bool fast_vote = false;    // the fp16 pass is worth trying on this data
  // fused tcgen05 sign vote (kernels_vote_mma.cuh): xh transposed (K-major B operand) + TMA tensor maps
  // tcgen05 integer Gram (kernels_gram_mma.cuh): pre-scaled transposed fp64 copy, M-tile and output tables
  int2* XsT = nullptr;       // [Ppad][ldx]  x' = x~ 2^(23 - e_p) as {rint(x'), fp32 bits of the remainder}
  double* xunit = nullptr;   // [Ppad] 2^(e_p - 23)
  int64_t ldx = 0;
  bool gram_mma = false;
  int4* gm_tiles = nullptr;  // [gm_n_tiles]
  void* gm_outs = nullptr;   // GramOut [gm_n_outs]
  int gm_n_tiles = 0, gm_n_outs = 0;
  const plspm_model* gm_model = nullptr;
  uint8_t* xt_img = nullptr;  // XhT tile images [chunk of 64 rows][column chunk of 256] 32 KB (vote_xt_image_kernel)
  uint8_t* xl_img = nullptr;  // block-column images [chunk][K = 16 block] 2 KB (vote_xl_image_kernel)
  int* lv_blk = nullptr;      // [L] first K = 16 block of every LV
  int vm_n_blocks = 0;
  int64_t n_chunks64 = 0;     // chunks of 64 rows
  bool mma_vote = false;
  // exact integer digit planes of x~ for the tensor-core column sums (see digits_kernel)
  int8_t* D8 = nullptr;      // [I8_DIGITS * Ppad][Npad]
  double* dscale = nullptr;  // [Ppad] value of one unit of the least significant digit
  int64_t Npad = 0;
  bool i8_colsum = false;
  // ... and of the pair products z_i = x~_ip x~_iq of the model's Gram tile set (see zdigits_kernel)
  int8_t* Z8 = nullptr;      // [I8_DIGITS * n_zcols][Npad] resident planes, or null: generated per row chunk per batch
  int *zp = nullptr, *zq = nullptr;       // [n_zcols] the two columns of every pair
  double* zqscale = nullptr;              // [n_zcols] 2^(40 - e_p - e_q)
  int64_t z_chunk_rows = 0;               // streaming mode: rows per generated chunk
  double* zdscale = nullptr; // [n_zcols]
  int *zdst = nullptr, *zdst2 = nullptr;  // [n_zcols] offsets into a replicate's tile array (mirror or -1)
  int n_zcols = 0;
  const plspm_model* z_model = nullptr;   // the tile set the planes were built for
  cudaStream_t stream = nullptr;
  Workspace ws;          // grown on demand, reused across calls
  StageTimer timer;
  int sm_count = 148;
  ImputeCtx* imp = nullptr;
  struct ImagePrefetch* prefetch = nullptr;  // multiplicity images of the first batch, generated during the upload (plspm_bootstrap_host)
  const uint8_t *c8img_ovr = nullptr, *c8vote_ovr = nullptr;  // ... and what the kernels of that batch read instead of the workspace images
  size_t ws_off = 0;       // workspace of the batch being enqueued (plspm_bootstrap pipelines two)
  bool img_ready = false;  // the multiplicity images of the batch in flight were written by resample_images_kernel
  int max_smem = 227 * 1024;
};

template <class T>
static int upload_vec(plspm_model* m, const std::vector<T>& v, const T** out) {
  void* p = nullptr;
  size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
  CK(cudaMalloc(&p, bytes));
  m->dev_allocs.push_back(p);
  if (!v.empty()) CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  *out = (const T*)p;
  return 0;
}

#include "ptx_util.cuh"
#include "kernels_resample.cuh"
#include "kernels_upload.cuh"
#include "kernels_digits.cuh"
#include "kernels_vote.cuh"
#include "kernels_vote_mma.cuh"
#include "kernels_gram_mma.cuh"
#include "kernels_numstep.cuh"
#include "kernels_gram.cuh"
#include "kernels_solve.cuh"
#include "kernels_impute.cuh"

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) costs microseconds per call and a single fit launches five kernels
// that need it: remember, per device and kernel, the largest size already granted.
template <typename F>
static cudaError_t ensure_smem(F* func, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> granted;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  size_t& cur = granted[{dev, (const void*)func}];
  if (bytes <= cur) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

extern "C" {

int plspm_version(void) { return 100; }
const char* plspm_last_error(void) { return g_err.c_str(); }

int plspm_device_count(int32_t* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) { *count = 0; return fail(PLSPM_ERR_CUDA, cudaGetErrorString(e)); }
  *count = n;
  return PLSPM_OK;
}
int plspm_set_device(int32_t device) {
  CK(cudaSetDevice(device));
  return PLSPM_OK;
}

int plspm_model_create(int32_t L, const int32_t* block_sizes, const int8_t* modes, const int8_t* path, int32_t scaled,
                       int32_t tile_policy, plspm_model** out) {
  if (!block_sizes || !modes || !path || !out) return fail(PLSPM_ERR_INVALID, "null argument");
  plspm_model* m = new plspm_model();
  std::string err;
  if (build_model(L, block_sizes, modes, path, scaled, tile_policy, m->h, err)) {
    delete m;
    return fail(PLSPM_ERR_INVALID, err);
  }
  if ((size_t)m->h.Ppad * 8 > 64 * 1024) {
    delete m;
    return fail(PLSPM_ERR_UNSUPPORTED, "more than 8192 (padded) manifest variables");
  }
  CK(cudaGetDevice(&m->device));
  HostModel& h = m->h;
  ModelView v = h.host_view();
  int rc = 0;
  rc |= upload_vec(m, h.lv_off, &v.lv_off);
  rc |= upload_vec(m, h.lv_k, &v.lv_k);
  rc |= upload_vec(m, h.lv_mode, &v.lv_mode);
  rc |= upload_vec(m, h.col_lv, &v.col_lv);
  rc |= upload_vec(m, h.col_src, &v.col_src);
  rc |= upload_vec(m, h.path, &v.path);
  rc |= upload_vec(m, h.tile_sa, &v.tile_sa);
  rc |= upload_vec(m, h.tile_sb, &v.tile_sb);
  rc |= upload_vec(m, h.tile_of, &v.tile_of);
  rc |= upload_vec(m, h.lane_tile, &v.lane_tile);
  rc |= upload_vec(m, h.pair_l, &v.pair_l);
  rc |= upload_vec(m, h.pair_j, &v.pair_j);
  rc |= upload_vec(m, h.pair_voff, &v.pair_voff);
  rc |= upload_vec(m, h.lv_pair_begin, &v.lv_pair_begin);
  rc |= upload_vec(m, h.eff_from, &v.eff_from);
  rc |= upload_vec(m, h.eff_to, &v.eff_to);
  rc |= upload_vec(m, h.chol_b_off, &v.chol_b_off);
  rc |= upload_vec(m, h.pred_begin, &v.pred_begin);
  rc |= upload_vec(m, h.pred_idx, &v.pred_idx);
  rc |= upload_vec(m, h.succ_begin, &v.succ_begin);
  rc |= upload_vec(m, h.succ_idx, &v.succ_idx);
  rc |= upload_vec(m, h.omega, &v.omega);
  if (rc) {
    plspm_model_destroy(m);
    return PLSPM_ERR_CUDA;
  }
  m->dv = v;
  *out = m;
  return PLSPM_OK;
}

void plspm_model_destroy(plspm_model* m) {
  if (!m) return;
  for (void* p : m->dev_allocs) cudaFree(p);
  delete m;
}

int plspm_model_query(const plspm_model* m, int32_t* info) {
  if (!m || !info) return fail(PLSPM_ERR_INVALID, "null argument");
  std::memset(info, 0, 16 * sizeof(int32_t));
  const HostModel& h = m->h;
  info[0] = h.L; info[1] = h.P; info[2] = h.Ppad; info[3] = h.n_tiles; info[4] = h.n_tg; info[5] = h.n_pairs;
  info[6] = h.n_eff; info[7] = h.n_out(); info[8] = h.full; info[9] = h.scaled; info[10] = h.n_cross; info[11] = m->numeric ? 1 : 0;
  int nz = 0;  // pair-product columns of the tile set (rows / 6 of the int8 Gram GEMM)
  for (int t = 0; t < h.n_tiles; ++t)
    for (int r = 0; r < SLOT; ++r)
      for (int c = (h.tile_sa[t] == h.tile_sb[t] ? r : 0); c < SLOT; ++c)
        nz += (h.col_lv[h.tile_sa[t] * SLOT + r] >= 0 && h.col_lv[h.tile_sb[t] * SLOT + c] >= 0) ? 1 : 0;
  info[12] = nz;
  return PLSPM_OK;
}

int plspm_model_set_numeric(plspm_model* m, int32_t on) {
  if (!m) return fail(PLSPM_ERR_INVALID, "null argument");
  m->numeric = on != 0;
  return PLSPM_OK;
}

int plspm_model_effects(const plspm_model* m, int32_t* from, int32_t* to) {
  if (!m || !from || !to) return fail(PLSPM_ERR_INVALID, "null argument");
  for (int e = 0; e < m->h.n_eff; ++e) { from[e] = m->h.eff_from[e]; to[e] = m->h.eff_to[e]; }
  return PLSPM_OK;
}

// A data handle only depends on the column layout of the model it was created with (block sizes in
// path order); any model with the same layout -- other modes, paths, tile policy, treatment -- may use it.
static bool same_layout(const plspm_model* a, const plspm_model* b) {
  if (a == b) return true;
  return a && b && a->device == b->device && a->h.L == b->h.L && a->h.P == b->h.P && a->h.Ppad == b->h.Ppad &&
         a->h.lv_off == b->h.lv_off && a->h.lv_k == b->h.lv_k && a->h.col_src == b->h.col_src;
}

static int ws_reserve(plspm_data* d, size_t bytes) {
  if (d->ws.bytes >= bytes) return 0;
  if (d->ws.ptr) g_pool.release(d->ws.ptr);
  d->ws.ptr = nullptr;
  d->ws.bytes = 0;
  CK(g_pool.alloc(&d->ws.ptr, bytes));
  d->ws.bytes = bytes;
  return 0;
}

int plspm_data_create(const plspm_model* m, const double* X, int64_t N, int64_t ld, int32_t x_is_device,
                      plspm_data** out) {
  if (!m || !X || !out) return fail(PLSPM_ERR_INVALID, "null argument");
  const HostModel& h = m->h;
  if (N < 2) return fail(PLSPM_ERR_INVALID, "need at least two observations");
  if (N >= ((int64_t)1 << 32)) return fail(PLSPM_ERR_UNSUPPORTED, "more than 2^32-1 observations");
  if (ld < h.P) return fail(PLSPM_ERR_INVALID, "leading dimension smaller than the number of manifest variables");
  plspm_data* d = new plspm_data();
  d->model = m;
  d->N = N;
  int dev = 0;
  // (two attribute queries, not cudaGetDeviceProperties: that call costs 2.5 ms -- and now and then 40 ms -- and
  //  plspm_bootstrap_host creates a data handle per call)
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&d->max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) {
    delete d;
    return fail(PLSPM_ERR_CUDA, "no CUDA device: plspm_b200 has no CPU path");
  }
  auto bail = [&](int rc) { plspm_data_destroy(d); return rc; };
  static const bool tracing = getenv("PLSPM_TRACE") != nullptr;
  const auto t_begin = std::chrono::steady_clock::now();
  auto trace = [&](const char* what) {
    if (!tracing) return;
    cudaStreamSynchronize(d->stream);
    fprintf(stderr, "[plspm_data_create] %-18s %8.3f ms\n", what,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
  };
  if (cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking) != cudaSuccess)
    return bail(fail(PLSPM_ERR_CUDA, "cudaStreamCreate failed"));
  cudaStream_t st = d->stream;
  trace("stream created");
  double* raw = nullptr;
  const double* Xd = X;
  int rc = 0;
  // scratch released (and pageable host staging kept alive) until the stream has drained at the end of the upload:
  // no intermediate synchronisation just to free a buffer
  std::vector<void*> deferred;
  std::vector<int> blk_col, lv_blk_host;
  auto body = [&]() -> int {
    CK(g_pool.alloc((void**)&d->X, (size_t)N * h.Ppad * sizeof(double)));
    CK(g_pool.alloc((void**)&d->mu, (size_t)h.Ppad * sizeof(double)));
    CK(cudaMemsetAsync(d->mu, 0, (size_t)h.Ppad * sizeof(double), st));
    if (!x_is_device) {
      CK(g_pool.alloc((void**)&raw, (size_t)N * h.P * sizeof(double)));
      if (ld == h.P)
        CK(cudaMemcpyAsync(raw, X, (size_t)N * h.P * sizeof(double), cudaMemcpyHostToDevice, st));
      else
        CK(cudaMemcpy2DAsync(raw, (size_t)h.P * sizeof(double), X, (size_t)ld * sizeof(double),
                             (size_t)h.P * sizeof(double), (size_t)N, cudaMemcpyHostToDevice, st));
      trace("h2d issued");
      Xd = raw;
      ld = h.P;
    }
    const int nblocks = (int)std::min<int64_t>(1024, (N + 63) / 64);
    const int64_t rpb = (N + nblocks - 1) / nblocks;
    double* partial = nullptr;
    CK(g_pool.alloc((void**)&partial, (size_t)nblocks * h.P * sizeof(double)));
    int* src_col = nullptr;
    CK(g_pool.alloc((void**)&src_col, (size_t)h.P * sizeof(int)));
    CK(cudaMemcpyAsync(src_col, h.src_col.data(), (size_t)h.P * sizeof(int), cudaMemcpyHostToDevice, st));
    d->timer.begin(ST_UPLOAD, st);
    colsum_partial_kernel<<<nblocks, 256, 0, st>>>(Xd, N, ld, h.P, rpb, partial);
    d->timer.end(st);
    d->timer.begin(ST_UPLOAD, st);
    colmean_final_kernel<<<(h.P + 3) / 4, 128, 0, st>>>(partial, nblocks, h.P, N, src_col, d->mu);
    d->timer.end(st);
    d->timer.begin(ST_UPLOAD, st);
    relayout_kernel<<<d->sm_count * 8, 256, 0, st>>>(Xd, N, ld, h.Ppad, m->dv.col_src, d->mu, d->X);
    d->timer.end(st);
    CK(cudaGetLastError());
    double* partial0 = nullptr;
    CK(g_pool.alloc((void**)&partial0, (size_t)nblocks * h.Ppad * sizeof(double)));
    CK(g_pool.alloc((void**)&d->colsum0, (size_t)h.Ppad * sizeof(double)));
    d->timer.begin(ST_UPLOAD, st);
    colsum_partial_kernel<<<nblocks, 256, 0, st>>>(d->X, N, h.Ppad, h.Ppad, rpb, partial0);
    d->timer.end(st);
    d->timer.begin(ST_UPLOAD, st);
    colsum_final_kernel<<<(h.Ppad + 3) / 4, 128, 0, st>>>(partial0, nblocks, h.Ppad, d->colsum0);
    d->timer.end(st);
    CK(cudaGetLastError());
    trace("relayout done");
    // fp16 copy for the tensor-core sign vote of sparse tile sets (PLSPM_VOTE=exact disables it)
    static const bool vote_exact = getenv("PLSPM_VOTE") && std::string(getenv("PLSPM_VOTE")) == "exact";
    int nsl_pad_chk = 1;
    while (nsl_pad_chk * SLOT < h.kmax) nsl_pad_chk <<= 1;
    static const bool vote_legacy_chk = getenv("PLSPM_VOTE") && std::string(getenv("PLSPM_VOTE")) == "cublas";
    const bool legacy_ok = nsl_pad_chk <= 32 && h.L * nsl_pad_chk <= SG_THREADS;
    const bool fused_ok = !vote_legacy_chk && h.kmax <= 16 * VM_MAX_K16;
    if (!h.full && !vote_exact && N >= 4096 && (fused_ok || legacy_ok)) {
      double* sq = nullptr;
      CK(g_pool.alloc((void**)&sq, (size_t)nblocks * h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->inv_sd, (size_t)h.Ppad * sizeof(double)));
      d->timer.begin(ST_UPLOAD, st);
      colsq_partial_kernel<<<nblocks, 256, 0, st>>>(d->X, N, h.Ppad, rpb, sq);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      inv_sd_kernel<<<(h.Ppad + 3) / 4, 128, 0, st>>>(sq, nblocks, h.Ppad, N, d->inv_sd);
      d->timer.end(st);
      // fused tcgen05 vote (default): blocks of at most 64 manifest variables (four K = 16 score steps)
      if (fused_ok && N < ((int64_t)1 << 31) - 256) {
        // operand images of the fused vote kernel (kernels_vote_mma.cuh): what its producer fetches with linear bulk copies
        d->n_chunks64 = (N + 63) / 64;
        const int n_pchunks = (h.Ppad + 255) / 256;
        std::vector<int>& lv_blk = lv_blk_host;
        lv_blk.assign(h.L, 0);
        for (int l = 0; l < h.L; ++l) {
          lv_blk[l] = (int)blk_col.size();
          for (int q = 0; q < (h.lv_k[l] + 15) / 16; ++q) blk_col.push_back(h.lv_off[l] + 16 * q);
        }
        d->vm_n_blocks = (int)blk_col.size();
        int* blk_col_dev = nullptr;
        CK(g_pool.alloc((void**)&blk_col_dev, blk_col.size() * sizeof(int)));
        CK(g_pool.alloc((void**)&d->lv_blk, lv_blk.size() * sizeof(int)));
        CK(cudaMemcpyAsync(blk_col_dev, blk_col.data(), blk_col.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->lv_blk, lv_blk.data(), lv_blk.size() * sizeof(int), cudaMemcpyHostToDevice, st));
        CK(g_pool.alloc((void**)&d->xt_img, (size_t)d->n_chunks64 * n_pchunks * VM_XT_BYTES));
        CK(g_pool.alloc((void**)&d->xl_img, (size_t)d->n_chunks64 * d->vm_n_blocks * VM_XL_BYTES));
        d->timer.begin(ST_UPLOAD, st);
        vote_xt_image_kernel<<<d->sm_count * 16, 256, 0, st>>>(d->X, N, h.Ppad, n_pchunks, d->n_chunks64, d->inv_sd, d->xt_img);
        d->timer.end(st);
        d->timer.begin(ST_UPLOAD, st);
        vote_xl_image_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N, h.Ppad, d->vm_n_blocks, d->n_chunks64, blk_col_dev, d->inv_sd,
                                                              d->xl_img);
        d->timer.end(st);
        CK(cudaGetLastError());
        deferred.push_back(blk_col_dev);
        d->mma_vote = true;
      } else {  // legacy route: fp32 score generation + library fp16 GEMM
        CK(g_pool.alloc((void**)&d->Xh, (size_t)N * h.Ppad * sizeof(__half)));
        d->timer.begin(ST_UPLOAD, st);
        make_half_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N, h.Ppad, d->inv_sd, d->Xh);
        d->timer.end(st);
        CK(g_pool.alloc((void**)&d->Xf, (size_t)N * h.Ppad * sizeof(float)));
        d->timer.begin(ST_UPLOAD, st);
        make_float_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N * h.Ppad, d->Xf);
        d->timer.end(st);
        CK(cudaGetLastError());
        d->blas = blas_acquire(dev);  // cublasCreate costs milliseconds: handles are recycled per device
        if (!d->blas) return fail(PLSPM_ERR_CUDA, "cublasCreate failed");
        d->blas_device = dev;
        cublasSetStream(d->blas, st);
      }
      deferred.push_back(sq);
      d->fast_vote = true;
      trace("fp16 copies");
    }
    static const bool conv_f32_env = getenv("PLSPM_CONV_F32") && atoi(getenv("PLSPM_CONV_F32")) != 0;
    if (!d->Xf && N >= 4096 && conv_f32_env) {  // fp32 copy for the (opt-in) fp32-screened non-metric criterion pass
      CK(g_pool.alloc((void**)&d->Xf, (size_t)N * h.Ppad * sizeof(float)));
      d->timer.begin(ST_UPLOAD, st);
      make_float_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N * h.Ppad, d->Xf);
      d->timer.end(st);
      CK(cudaGetLastError());
    }
    // tcgen05 integer Gram: pre-scaled transposed copy x' (digits are generated on the fly, no resident planes).
    // PLSPM_GRAM=cublas selects round 1's digit planes + library GEMM, PLSPM_GRAM=fp64 the fp64 kernels.
    static const std::string gram_env = getenv("PLSPM_GRAM") ? getenv("PLSPM_GRAM") : "";
    bool gram_heavy_tail = false;
    if (gram_env != "cublas" && gram_env != "fp64" && N >= 4096 && N < ((int64_t)1 << 31) - 256) {
      d->Npad = (N + 15) / 16 * 16;
      d->ldx = (N + 31) / 32 * 32;
      double *amax = nullptr, *xscale = nullptr;
      CK(g_pool.alloc((void**)&amax, (size_t)nblocks * h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&xscale, (size_t)h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->xunit, (size_t)h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->XsT, (size_t)h.Ppad * d->ldx * sizeof(int2)));
      d->timer.begin(ST_UPLOAD, st);
      colabsmax_partial_kernel<<<nblocks, 256, 0, st>>>(d->X, N, h.Ppad, rpb, amax);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      gram_xunit_kernel<<<(h.Ppad + 3) / 4, 128, 0, st>>>(amax, nblocks, h.Ppad, d->xunit, xscale);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      gram_xst_kernel<<<dim3((unsigned)((d->ldx + 31) / 32), (h.Ppad + 31) / 32), 256, 0, st>>>(d->X, N, h.Ppad, d->ldx, xscale,
                                                                                            d->XsT);
      d->timer.end(st);
      // heavy-tail guard (kernels_gram_mma.cuh): columns whose bound is set by <= 8 gross outliers
      int *tail_part = nullptr, *tail = nullptr;
      CK(g_pool.alloc((void**)&tail_part, (size_t)nblocks * h.Ppad * sizeof(int)));
      CK(g_pool.alloc((void**)&tail, (size_t)h.Ppad * sizeof(int)));
      d->timer.begin(ST_UPLOAD, st);
      gram_tailcount_partial_kernel<<<nblocks, 256, 0, st>>>(d->X, N, h.Ppad, rpb, d->xunit, tail_part);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      gram_tailcount_final_kernel<<<(h.Ppad + 3) / 4, 128, 0, st>>>(tail_part, nblocks, h.Ppad, tail);
      d->timer.end(st);
      CK(cudaGetLastError());
      std::vector<int> tail_host(h.Ppad);
      CK(cudaMemcpyAsync(tail_host.data(), tail, (size_t)h.Ppad * sizeof(int), cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      g_pool.release(amax);
      g_pool.release(xscale);
      g_pool.release(tail_part);
      g_pool.release(tail);
      bool heavy = false;
      for (int p = 0; p < h.Ppad; ++p) heavy |= h.col_lv[p] >= 0 && tail_host[p] >= 1 && tail_host[p] <= 8;
      if (heavy) {  // fp64 kernels for this data (second moments and column sums); the sign vote falls back too
        g_pool.release(d->XsT);
        g_pool.release(d->xunit);
        d->XsT = nullptr; d->xunit = nullptr;
        trace("heavy-tailed column: fp64 second moments");
      } else {
        d->gram_mma = true;
        d->i8_colsum = true;  // (the int8 multiplicities of a batch are built for this route)
        trace("pre-scaled transposed copy");
      }
      gram_heavy_tail = heavy;
    }
    // integer digit planes for the tensor-core column sums (PLSPM_COLSUM=fp64 keeps the fp64 kernel)
    static const bool colsum_fp64 = (getenv("PLSPM_COLSUM") && std::string(getenv("PLSPM_COLSUM")) == "fp64") || gram_env == "fp64";
    if (!d->gram_mma && !gram_heavy_tail && !colsum_fp64 && N >= 4096 && N < ((int64_t)1 << 31) - 16) {
      if (!d->blas) {
        d->blas = blas_acquire(dev);
        if (!d->blas) return fail(PLSPM_ERR_CUDA, "cublasCreate failed");
        d->blas_device = dev;
        cublasSetStream(d->blas, st);
      }
      d->Npad = (N + 15) / 16 * 16;
      double *amax = nullptr, *qscale = nullptr;
      CK(g_pool.alloc((void**)&amax, (size_t)nblocks * h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&qscale, (size_t)h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->dscale, (size_t)h.Ppad * sizeof(double)));
      CK(g_pool.alloc((void**)&d->D8, (size_t)I8_DIGITS * h.Ppad * d->Npad));
      d->timer.begin(ST_UPLOAD, st);
      colabsmax_partial_kernel<<<nblocks, 256, 0, st>>>(d->X, N, h.Ppad, rpb, amax);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      digit_scale_kernel<<<(h.Ppad + 127) / 128, 128, 0, st>>>(amax, nblocks, h.Ppad, d->dscale, qscale);
      d->timer.end(st);
      d->timer.begin(ST_UPLOAD, st);
      digits_kernel<<<dim3((h.Ppad + 31) / 32, (unsigned)((d->Npad + 127) / 128)), 256, 0, st>>>(d->X, N, h.Ppad, d->Npad,
                                                                                            qscale, d->D8);
      d->timer.end(st);
      CK(cudaGetLastError());
      CK(cudaStreamSynchronize(st));
      g_pool.release(amax);
      d->i8_colsum = true;
      trace("digit planes");
      // pair-product planes of the model's tile set, if they fit the budget (PLSPM_I8_GRAM_GB, default 24)
      static const double z_budget_gb = getenv("PLSPM_I8_GRAM_GB") ? atof(getenv("PLSPM_I8_GRAM_GB")) : 24.0;
      std::vector<int> zp, zq, zd1, zd2;
      for (int t = 0; t < h.n_tiles; ++t) {
        const int sa = h.tile_sa[t], sb = h.tile_sb[t];
        for (int r = 0; r < SLOT; ++r)
          for (int c = (sa == sb ? r : 0); c < SLOT; ++c) {
            const int pp = sa * SLOT + r, qq = sb * SLOT + c;
            if (h.col_lv[pp] < 0 || h.col_lv[qq] < 0) continue;  // padding columns are zero: their moments stay 0
            zp.push_back(pp); zq.push_back(qq);
            zd1.push_back(t * TILE + r * SLOT + c);
            zd2.push_back(sa == sb && r != c ? t * TILE + c * SLOT + r : -1);
          }
      }
      const double z_bytes = (double)I8_DIGITS * zp.size() * d->Npad;
      static const double z_chunk_gb = getenv("PLSPM_I8_CHUNK_GB") ? atof(getenv("PLSPM_I8_CHUNK_GB")) : 12.0;
      if (!zp.empty() && z_budget_gb > 0 && (int64_t)I8_DIGITS * (int64_t)zp.size() < ((int64_t)1 << 31)) {
        const int nz = (int)zp.size();
        CK(g_pool.alloc((void**)&d->zp, (size_t)nz * 4));
        CK(g_pool.alloc((void**)&d->zq, (size_t)nz * 4));
        CK(g_pool.alloc((void**)&d->zdst, (size_t)nz * 4));
        CK(g_pool.alloc((void**)&d->zdst2, (size_t)nz * 4));
        CK(g_pool.alloc((void**)&d->zqscale, (size_t)nz * 8));
        CK(g_pool.alloc((void**)&d->zdscale, (size_t)nz * 8));
        CK(cudaMemcpyAsync(d->zp, zp.data(), (size_t)nz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->zq, zq.data(), (size_t)nz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->zdst, zd1.data(), (size_t)nz * 4, cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(d->zdst2, zd2.data(), (size_t)nz * 4, cudaMemcpyHostToDevice, st));
        d->timer.begin(ST_UPLOAD, st);
        zscale_kernel<<<(nz + 127) / 128, 128, 0, st>>>(nz, d->zp, d->zq, qscale, d->dscale, d->zqscale, d->zdscale);
        d->timer.end(st);
        CK(cudaGetLastError());
        if (z_bytes <= z_budget_gb * 1e9) {  // resident planes: generated once
          CK(g_pool.alloc((void**)&d->Z8, (size_t)I8_DIGITS * nz * d->Npad));
          d->timer.begin(ST_UPLOAD, st);
          zdigits_kernel<<<dim3((nz + 31) / 32, (unsigned)((d->Npad + 127) / 128)), 256, 0, st>>>(
              d->X, N, h.Ppad, 0, d->Npad, nz, d->zp, d->zq, d->zqscale, d->Z8);
          d->timer.end(st);
          CK(cudaGetLastError());
        } else {  // too large to keep: a chunk of rows is generated per batch, right before its GEMM
          int64_t rows = (int64_t)(z_chunk_gb * 1e9 / ((double)I8_DIGITS * nz));
          rows = std::min<int64_t>(rows / 128 * 128, I8_KCHUNK);
          d->z_chunk_rows = std::max<int64_t>(rows, 4096);
        }
        CK(cudaStreamSynchronize(st));  // (the host vectors above are pageable: copies are done by now)
        d->n_zcols = nz;
        d->z_model = m;
        trace("pair-product planes");
      }
      g_pool.release(qscale);
    }
    CK(cudaStreamSynchronize(st));
    d->timer.collect();
    trace("timer collected");
    {  // a NaN / Inf anywhere in a column poisons its mean: one P-length check instead of a host pass over X
      std::vector<double> mu_host(h.Ppad);
      CK(cudaMemcpy(mu_host.data(), d->mu, (size_t)h.Ppad * sizeof(double), cudaMemcpyDeviceToHost));
      for (double v : mu_host)
        if (!std::isfinite(v)) return fail(PLSPM_ERR_INVALID, "non-finite values in the observation matrix");
    }
    g_pool.release(partial);
    g_pool.release(partial0);
    g_pool.release(src_col);
    for (void* q : deferred) g_pool.release(q);
    deferred.clear();
    return 0;
  };
  rc = body();
  if (!deferred.empty()) {  // (an error path left them behind)
    cudaStreamSynchronize(d->stream);
    for (void* q : deferred) g_pool.release(q);
  }
  if (raw) g_pool.release(raw);
  if (rc) return bail(rc);
  *out = d;
  return PLSPM_OK;
}

void plspm_data_destroy(plspm_data* d) {
  if (!d) return;
  if (d->X) g_pool.release(d->X);
  if (d->mu) g_pool.release(d->mu);
  if (d->colsum0) g_pool.release(d->colsum0);
  if (d->imp) {
    for (void* q : {(void*)d->imp->ax, (void*)d->imp->am, (void*)d->imp->mu_base, (void*)d->imp->G, (void*)d->imp->colsum, (void*)d->imp->ws})
      if (q) g_pool.release(q);
    delete d->imp;
  }
  if (d->Xh) g_pool.release(d->Xh);
  if (d->xt_img) g_pool.release(d->xt_img);
  if (d->xl_img) g_pool.release(d->xl_img);
  if (d->lv_blk) g_pool.release(d->lv_blk);
  if (d->XsT) g_pool.release(d->XsT);
  if (d->xunit) g_pool.release(d->xunit);
  if (d->gm_tiles) g_pool.release(d->gm_tiles);
  if (d->gm_outs) g_pool.release(d->gm_outs);
  if (d->Xf) g_pool.release(d->Xf);
  if (d->inv_sd) g_pool.release(d->inv_sd);
  if (d->D8) g_pool.release(d->D8);
  if (d->Z8) g_pool.release(d->Z8);
  if (d->zdscale) g_pool.release(d->zdscale);
  if (d->zp) g_pool.release(d->zp);
  if (d->zq) g_pool.release(d->zq);
  if (d->zqscale) g_pool.release(d->zqscale);
  if (d->zdst) g_pool.release(d->zdst);
  if (d->zdst2) g_pool.release(d->zdst2);
  if (d->dscale) g_pool.release(d->dscale);
  if (d->blas) blas_release(d->blas_device, d->blas);
  if (d->ws.ptr) g_pool.release(d->ws.ptr);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
}

// ------------------------------------------------------------------------------------------------
// launch plans, workspace layout, batch driver
// ------------------------------------------------------------------------------------------------
struct StreamPlan {  // one streaming pass over X (Gram tiles or cross-moment tiles)
  int RT = 0, stages = 0, n_chunks = 1;
  int64_t chunk_rows = 0, n_groups = 0;
  size_t smem = 0;
};
struct BatchPlan {
  StreamPlan gram, cross;
  int gm_ksplit = 1, gm_rows = 0, gm_groups = 1;  // tcgen05 integer Gram: row ranges, rows per range, 512-replicate groups
  int64_t gm_nb_pad = 0;
  int cs_chunks = 1;  // row chunks of the column-sum kernel
  int64_t cs_chunk_rows = 0;
  // criterion pass of the numeric non-metric path
  int cv_nsl_pad = 1, cv_reps_per_cta = 1, cv_rows = 1;
  bool cv_f32 = false;
  unsigned cv_gx = 1, cv_gy = 1;
  size_t cv_smem = 0;
};

static void plan_colsum(const plspm_data* d, int64_t nb, BatchPlan& g) {
  const HostModel& h = d->model->h;
  const int64_t blocks_xy = (int64_t)((h.Ppad + CS_COLS - 1) / CS_COLS) * ((nb + CS_REPS - 1) / CS_REPS);
  const int64_t want = (int64_t)d->sm_count * 8;
  int64_t chunks = std::max<int64_t>(1, (want + blocks_xy - 1) / blocks_xy);
  chunks = std::min<int64_t>(chunks, std::max<int64_t>(1, d->N / (CS_ROWS * 8)));
  chunks = std::min<int64_t>(chunks, 65535);
  int64_t rows = ((d->N + chunks - 1) / chunks + CS_ROWS - 1) / CS_ROWS * CS_ROWS;
  g.cs_chunk_rows = rows;
  g.cs_chunks = (int)((d->N + rows - 1) / rows);
}

// Ring geometry + row chunking of a streaming pass.  extra_row_bytes / extra_fixed_bytes: per-CTA
// shared memory the kernel needs per ring row and in total besides the ring (cross mode scratch).
static int plan_stream(const plspm_data* d, int64_t n_items, size_t extra_row_bytes, size_t extra_fixed_bytes,
                       StreamPlan& g) {
  const HostModel& h = d->model->h;
  const size_t row_bytes = (size_t)h.Ppad * 8;
  // stages of <= 64 KB / 32 rows; measured on c3: 32-row stages beat 16-row stages by 10 % (per-tile
  // handshakes amortise), three stages are enough since the last consumer of a stage refills it at once
  static const int stage_kb = getenv("PLSPM_GRAM_STAGE_KB") ? atoi(getenv("PLSPM_GRAM_STAGE_KB")) : 64;
  static const int want_stages = getenv("PLSPM_GRAM_STAGES") ? atoi(getenv("PLSPM_GRAM_STAGES")) : 3;
  const size_t budget = (size_t)d->max_smem - 8 * 1024;  // static shared memory (barriers, row lists) + slack
  if (budget < extra_fixed_bytes + 2 * (row_bytes + extra_row_bytes))
    return fail(PLSPM_ERR_UNSUPPORTED, "manifest rows too wide for the shared-memory ring");
  // Rows per stage matter more than the number of stages (every tile costs a list build, a pipeline
  // prologue/epilogue and, for odd counts, one padded row): measured on P=1024, 2 stages x 12 rows beat
  // 3 x 8 by 14 % and 6 x 4 by 41 %.  So: three stages if they still hold >= 16 rows each, else two.
  static const bool stage_env = getenv("PLSPM_GRAM_STAGE_KB") != nullptr;
  auto rows_for = [&](int st) {
    int64_t r = (int64_t)((budget - extra_fixed_bytes) / ((size_t)st * row_bytes + extra_row_bytes));
    if (stage_env || st >= 3) r = std::min<int64_t>(r, std::max<int64_t>(1, (int64_t)((size_t)stage_kb * 1024 / row_bytes)));
    return std::min<int64_t>(r, 32);
  };
  int stages = std::max(2, std::min(want_stages, GRAM_MAX_STAGES));
  int64_t RT = rows_for(stages);
  if (stages > 2 && RT < 16 && rows_for(2) > RT) {
    stages = 2;
    RT = rows_for(2);
  }
  if (RT < 1) return fail(PLSPM_ERR_UNSUPPORTED, "manifest rows too wide for the shared-memory ring");
  const int64_t n_tiles_rt = (d->N + RT - 1) / RT;
  stages = (int)std::min<int64_t>(stages, std::max<int64_t>(2, n_tiles_rt));
  const int64_t n_groups = (n_items + GRAM_WARPS - 1) / GRAM_WARPS;
  // split rows only when there are too few (replicate, tile group) items to fill the chip
  int64_t n_chunks = 1;
  const int64_t want = (int64_t)d->sm_count * 2;
  if (n_groups < want) n_chunks = (want + n_groups - 1) / n_groups;
  n_chunks = std::max<int64_t>(1, std::min<int64_t>(n_chunks, n_tiles_rt / 4 > 0 ? n_tiles_rt / 4 : 1));
  int64_t chunk_rows = ((n_tiles_rt + n_chunks - 1) / n_chunks) * RT;
  n_chunks = (d->N + chunk_rows - 1) / chunk_rows;
  g.RT = (int)RT; g.stages = stages; g.n_chunks = (int)n_chunks; g.chunk_rows = chunk_rows; g.n_groups = n_groups;
  g.smem = (size_t)stages * RT * row_bytes + extra_fixed_bytes + (size_t)RT * extra_row_bytes;
  return 0;
}

// M tiles and output moments of the tcgen05 integer Gram for the model's tile set (kernels_gram_mma.cuh).
static int gm_build_tables(plspm_data* d) {
  const plspm_model* m = d->model;
  if (d->gm_model == m) return 0;
  const HostModel& h = m->h;
  std::vector<int4> tiles;
  std::vector<GramOut> outs;
  auto add_pair = [&](int mt_base, int j, int p, int q, int dst1, int dst2) {
    const int T0 = (6 * j) / 128, T1 = (6 * j + 5) / 128;
    GramOut go;
    go.slot0 = (mt_base + T0) * GM_PAIRS_PER_TILE + (j - 21 * T0);
    go.slot1 = T1 != T0 ? (mt_base + T1) * GM_PAIRS_PER_TILE + (j - 21 * T1) : -1;
    go.p = p; go.q = q; go.dst1 = dst1; go.dst2 = dst2;
    outs.push_back(go);
  };
  for (int t = 0; t < h.n_tiles; ++t) {
    const int sa = h.tile_sa[t], sb = h.tile_sb[t];
    const int base = (int)tiles.size();
    if (sa != sb) {
      for (int T = 0; T < 3; ++T) tiles.push_back(make_int4(GM_KIND_OFF, T, sa, sb));
      for (int r = 0; r < SLOT; ++r)
        for (int c = 0; c < SLOT; ++c) {
          const int p = sa * SLOT + r, q = sb * SLOT + c;
          if (h.col_lv[p] < 0 || h.col_lv[q] < 0) continue;  // padding columns: their moments stay 0
          add_pair(base, r * SLOT + c, p, q, t * TILE + r * SLOT + c, -1);
        }
    } else {
      for (int T = 0; T < 2; ++T) tiles.push_back(make_int4(GM_KIND_DIAG, T, sa, sa));
      int j = 0;
      for (int r = 0; r < SLOT; ++r)
        for (int c = r; c < SLOT; ++c, ++j) {
          const int p = sa * SLOT + r, q = sa * SLOT + c;
          if (h.col_lv[p] < 0 || h.col_lv[q] < 0) continue;
          add_pair(base, j, p, q, t * TILE + r * SLOT + c, r != c ? t * TILE + c * SLOT + r : -1);
        }
    }
  }
  for (int sa = 0; sa < h.ns; sa += 2) {  // column sums, two slots per M tile
    const int sb = sa + 1 < h.ns ? sa + 1 : sa;
    const int base = (int)tiles.size();
    tiles.push_back(make_int4(GM_KIND_SUM2, 0, sa, sb));
    for (int j = 0; j < 2 * SLOT; ++j) {
      if (j >= SLOT && sb == sa) break;
      const int p = (j < SLOT ? sa : sb) * SLOT + (j & 7);
      if (h.col_lv[p] < 0) continue;
      add_pair(base, j, p, -1, p, -1);
    }
  }
  while (tiles.size() % 4) tiles.push_back(make_int4(GM_KIND_NONE, 0, 0, 0));  // whole clusters of up to 4 M tiles
  if (d->gm_tiles) g_pool.release(d->gm_tiles);
  if (d->gm_outs) g_pool.release(d->gm_outs);
  d->gm_tiles = nullptr; d->gm_outs = nullptr; d->gm_model = nullptr;
  CK(g_pool.alloc((void**)&d->gm_tiles, tiles.size() * sizeof(int4)));
  CK(g_pool.alloc((void**)&d->gm_outs, outs.size() * sizeof(GramOut)));
  CK(cudaMemcpyAsync(d->gm_tiles, tiles.data(), tiles.size() * sizeof(int4), cudaMemcpyHostToDevice, d->stream));
  CK(cudaMemcpyAsync(d->gm_outs, outs.data(), outs.size() * sizeof(GramOut), cudaMemcpyHostToDevice, d->stream));
  CK(cudaStreamSynchronize(d->stream));  // (pageable host vectors)
  d->gm_n_tiles = (int)tiles.size();
  d->gm_n_outs = (int)outs.size();
  d->gm_model = m;
  return 0;
}
static int gm_count_tiles(const HostModel& h) {
  int n = (h.ns + 1) / 2;
  for (int t = 0; t < h.n_tiles; ++t) n += h.tile_sa[t] != h.tile_sb[t] ? 3 : 2;
  return (n + 3) / 4 * 4;
}
// row ranges of the integer Gram: among the splits with at most 24 ranges the one whose CTA count fills whole waves best
static void gm_plan(const plspm_data* d, int64_t nb, BatchPlan& bp) {
  const int n_mtiles = gm_count_tiles(d->model->h);
  bp.gm_groups = (int)((nb + 511) / 512);
  bp.gm_nb_pad = (nb + 31) / 32 * 32;
  const int64_t tiles0 = (int64_t)n_mtiles * bp.gm_groups;
  double best = 1e300;
  for (int k = 1; k <= 24; ++k) {
    const int64_t rows = ((d->N + k - 1) / k + GM_STAGE_ROWS - 1) / GM_STAGE_ROWS * GM_STAGE_ROWS;
    if ((int64_t)k * rows - d->N >= rows && k > 1) continue;  // an empty last range
    // the partial moments of every range are kept until the recombination: at most 1 GB of them
    if (k > 1 && (double)k * n_mtiles * GM_PAIRS_PER_TILE * bp.gm_nb_pad * sizeof(longlong2) > 1.0e9) break;
    const double cost = (double)((tiles0 * k + d->sm_count - 1) / d->sm_count) * (double)(rows + 1024);
    if (cost < best * 0.97) { best = cost; bp.gm_ksplit = k; bp.gm_rows = (int)rows; }  // (prefer fewer partial buffers)
  }
}

static int plan_batch(const plspm_data* d, int64_t nb, BatchPlan& bp) {
  const HostModel& h = d->model->h;
  if (int rc = plan_stream(d, nb * h.n_tg, 0, 0, bp.gram)) return rc;
  if (!h.full && !d->model->numeric) {
    const size_t per_row = (size_t)GRAM_WARPS * h.ng * SLOT * 8;  // score scratch row of every warp
    if (int rc = plan_stream(d, nb * h.n_tg_cross, per_row, 8 * per_row, bp.cross)) return rc;
  }
  plan_colsum(d, nb, bp);
  if (d->gram_mma) gm_plan(d, nb, bp);
  if (d->model->numeric) {
    int nsl_pad = 1;
    while (nsl_pad * SLOT < h.kmax) nsl_pad <<= 1;
    if (nsl_pad > 32 || h.L * nsl_pad > SG_THREADS)
      return fail(PLSPM_ERR_UNSUPPORTED, "numeric non-metric path: L x (padded slots per block) exceeds 256");
    const int nbl = SG_THREADS / (h.L * nsl_pad);
    bp.cv_nsl_pad = nsl_pad;
    // the fp32-screened variant is correct but not yet faster than the fp64 one (12.6 vs ~10 ms per pass on c3:
    // branchy inner loop, 128-register cap); opt-in until it is tuned
    static const bool conv_f32 = getenv("PLSPM_CONV_F32") && atoi(getenv("PLSPM_CONV_F32")) != 0;
    bp.cv_f32 = conv_f32 && d->Xf != nullptr;
    const size_t esz = bp.cv_f32 ? 4 : 8;
    const int rpt = bp.cv_f32 ? CvTraits<float>::RPT : CvTraits<double>::RPT;
    bp.cv_reps_per_cta = nbl * rpt;
    const size_t fixed = (size_t)SG_THREADS * rpt * 8 + 16;
    // at most 64 rows per staged tile, and small enough for two CTAs per SM
    bp.cv_rows = (int)std::max<size_t>(1, std::min<size_t>(bp.cv_f32 ? 64 : 32, ((size_t)d->max_smem / 2 - 8192 - fixed) /
                                                                                 (((size_t)h.Ppad + bp.cv_reps_per_cta) * esz)));
    bp.cv_smem = (size_t)bp.cv_rows * ((size_t)h.Ppad + bp.cv_reps_per_cta) * esz + fixed;
    bp.cv_gy = (unsigned)((nb + bp.cv_reps_per_cta - 1) / bp.cv_reps_per_cta);
    bp.cv_gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((d->N + bp.cv_rows - 1) / bp.cv_rows,
                                                                 (4 * d->sm_count + bp.cv_gy - 1) / bp.cv_gy));
  }
  return 0;
}

static size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// Device workspace of one batch of up to nb replicates (offsets into plspm_data::ws).
constexpr int FAST_RC = 4096;  // rows per tensor-core GEMM chunk (bounds the fp32 accumulation error)
struct BatchBuffers {
  size_t total = 0;
  size_t counts, idx, G, Gpart, colsum, cspart, ws, out, iters, status, wf, CG, CGpart, sh, BT, Cf, rep_map, state;
  // single-fit outputs
  size_t weights, loadings, r2, paths, totalfx, crossl, coef, shift, scores;
  // numeric non-metric path: per-replicate iteration state
  size_t num_a, num_co, num_cn, num_so, num_sn, num_meta, num_done, num_cpart, num_cmain;
  // tensor-core column sums: int8 multiplicities, int32 digit sums, overflow flag
  size_t c8, c8img, c8vote, s32, ovf, zs32, zchunk, gm_part;
};
static BatchBuffers layout_batch(const plspm_data* d, int64_t nb, const BatchPlan& bp, bool with_counts, bool with_idx,
                                 bool rows_on_device_of_caller, bool single_fit, bool want_scores) {
  const HostModel& h = d->model->h;
  BatchBuffers b;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(std::max<size_t>(bytes, 8)); return o; };
  const size_t gsz = (size_t)h.n_tiles * TILE * 8, csz = (size_t)h.n_cross * TILE * 8;
  b.counts = take(with_counts ? (size_t)nb * d->N * 4 : 0);
  b.idx = take(with_idx ? (size_t)nb * d->N * 4 : 0);
  b.G = take((size_t)nb * gsz);
  b.Gpart = take(bp.gram.n_chunks > 1 ? (size_t)nb * gsz * bp.gram.n_chunks : 0);
  b.colsum = take((size_t)nb * h.Ppad * 8);
  b.cspart = take(bp.cs_chunks > 1 ? (size_t)nb * h.Ppad * 8 * bp.cs_chunks : 0);
  b.ws = take((size_t)nb * h.ws_doubles * 8);
  b.out = take(rows_on_device_of_caller || single_fit ? 0 : (size_t)nb * h.n_out() * 8);
  b.iters = take((size_t)nb * 4);
  b.status = take((size_t)nb * 4);
  const bool numeric = d->model->numeric;
  const bool vote = !h.full && !numeric;  // sparse tile set: buffers of the cross-moment / sign-vote pass
  b.wf = take(vote ? (size_t)nb * h.Ppad * 8 : 0);
  b.CG = take(vote ? (size_t)nb * csz : 0);
  b.CGpart = take(vote && bp.cross.n_chunks > 1 ? (size_t)nb * csz * bp.cross.n_chunks : 0);
  b.sh = take(vote ? (size_t)nb * h.L * 8 : 0);
  b.state = take(vote ? (size_t)nb * plspm::solver_state_doubles(d->model->dv) * 8 : 0);
  const bool fast = vote && d->fast_vote && !single_fit;
  b.num_a = take(numeric ? (size_t)nb * h.Ppad * 8 : 0);
  b.num_co = take(numeric ? (size_t)nb * h.Ppad * 8 : 0);
  b.num_cn = take(numeric ? (size_t)nb * h.Ppad * 8 : 0);
  b.num_so = take(numeric ? (size_t)nb * h.L * 8 : 0);
  b.num_sn = take(numeric ? (size_t)nb * h.L * 8 : 0);
  b.num_meta = take(numeric ? (size_t)nb * 16 : 0);
  b.num_done = take(8);
  const bool i8 = d->i8_colsum && with_counts;
  b.c8 = take(i8 && !d->gram_mma ? (size_t)nb * d->Npad : 0);  // plain [nb][Npad] (digit-plane GEMMs of round 1)
  // tile images of the multiplicities for the tensor-core kernels: [stage of 128 rows][group of 512 replicates] 64 KB
  b.c8img = take(i8 && d->gram_mma ? (size_t)((d->N + 127) / 128) * ((nb + 511) / 512) * 65536 : 0);
  // ... and for the fused sign vote: [chunk of 64 rows][tile of 128 replicates] 8 KB
  b.c8vote = take(i8 && d->mma_vote && !h.full && !d->model->numeric ? (size_t)((d->N + 63) / 64) * ((nb + 127) / 128) * 8192 : 0);
  b.s32 = take(i8 && !d->gram_mma ? (size_t)nb * I8_DIGITS * h.Ppad * sizeof(int32_t) : 0);
  b.gm_part = take(i8 && d->gram_mma ? (size_t)bp.gm_ksplit * gm_count_tiles(h) * GM_PAIRS_PER_TILE * bp.gm_nb_pad * sizeof(longlong2) : 0);
  b.ovf = take(8);
  b.zs32 = take(i8 && d->n_zcols ? (size_t)nb * I8_DIGITS * d->n_zcols * sizeof(int32_t) : 0);
  b.zchunk = take(i8 && d->n_zcols && !d->Z8 ? (size_t)I8_DIGITS * d->n_zcols * d->z_chunk_rows : 0);
  b.num_cpart = take(numeric ? (size_t)nb * bp.cv_gx * 8 : 0);
  b.num_cmain = take(numeric ? (size_t)nb * 8 : 0);
  const size_t ldl = (size_t)(nb + 7) / 8 * 8;  // replicate stride of the LV-major score / cross-moment layout
  b.BT = take(fast && !d->mma_vote ? ldl * h.L * FAST_RC * sizeof(__half) : 0);
  b.Cf = take(fast ? ldl * h.L * h.Ppad * sizeof(float) : 0);
  b.rep_map = take((size_t)nb * 4);
  if (single_fit) {
    const size_t L = h.L, P = h.P;
    b.weights = take(P * 8); b.loadings = take(P * 8); b.r2 = take(L * 8); b.paths = take(L * L * 8);
    b.totalfx = take(L * L * 8); b.crossl = take(P * L * 8); b.coef = take((size_t)h.Ppad * 8); b.shift = take(L * 8);
    b.scores = take(want_scores ? (size_t)d->N * L * 8 : 0);
  }
  b.total = off;
  return b;
}

static int launch_stream(plspm_data* d, bool cross, int64_t nb, const uint32_t* counts_dev, const StreamPlan& sp,
                         double* out_tiles, double* part_tiles, const double* wf, const int* rep_map = nullptr) {
  const plspm_model* m = d->model;
  const HostModel& h = m->h;
  cudaStream_t st = d->stream;
  GramParams p;
  p.X = d->X; p.counts = counts_dev; p.N = d->N; p.Ppad = h.Ppad;
  p.n_tiles = cross ? h.n_cross : h.n_tiles;
  p.n_tg = cross ? h.n_tg_cross : h.n_tg;
  p.tile_sa = m->dv.tile_sa; p.tile_sb = m->dv.tile_sb; p.lane_tile = m->dv.lane_tile;
  p.n_items = nb * p.n_tg; p.n_chunks = sp.n_chunks; p.chunk_rows = sp.chunk_rows; p.RT = sp.RT; p.stages = sp.stages;
  p.G = (sp.n_chunks > 1) ? part_tiles : out_tiles;
  p.L = h.L; p.ng = h.ng; p.lv_off = m->dv.lv_off; p.lv_k = m->dv.lv_k; p.wf = wf; p.rep_map = rep_map;
  const int64_t n_groups = (p.n_items + GRAM_WARPS - 1) / GRAM_WARPS;
  const int64_t grid = n_groups * sp.n_chunks;
  if (grid > 0x7fffffff) return fail(PLSPM_ERR_UNSUPPORTED, "batch too large for one launch");
  d->timer.begin(cross ? ST_CROSS : ST_GRAM, st);
  if (cross) {
    CK(ensure_smem(gram_kernel<true>, (size_t)(sp.smem)));
    gram_kernel<true><<<(unsigned)grid, GRAM_THREADS, sp.smem, st>>>(p);
  } else {
    CK(ensure_smem(gram_kernel<false>, (size_t)(sp.smem)));
    gram_kernel<false><<<(unsigned)grid, GRAM_THREADS, sp.smem, st>>>(p);
  }
  d->timer.end(st);
  CK(cudaGetLastError());
  if (sp.n_chunks > 1) {
    d->timer.begin(ST_REDUCE, st);
    reduce_chunks_kernel<<<d->sm_count * 4, 256, 0, st>>>(part_tiles, nb, sp.n_chunks, (int64_t)p.n_tiles * TILE,
                                                          out_tiles);
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  return 0;
}

// Exact redo (fp64 cross moments) of the replicates the low-precision vote could not decide.
static int redo_exact(plspm_data* d, int64_t n_list, const int* rep_map_dev, const uint32_t* counts_dev,
                      const BatchBuffers& bb, int scheme, double tol, int max_iter, const BatchPlan& bp,
                      double* out_rows) {
  const plspm_model* m = d->model;
  const HostModel& h = m->h;
  cudaStream_t st = d->stream;
  char* base = (char*)d->ws.ptr + d->ws_off;
  auto D = [&](size_t o) { return (double*)(base + o); };
  StreamPlan sp = bp.cross;
  sp.n_chunks = 1;  // partial buffers are laid out per replicate position; keep whole-row passes here
  sp.chunk_rows = ((d->N + sp.RT - 1) / sp.RT) * sp.RT;
  if (int rc = launch_stream(d, true, n_list, counts_dev, sp, D(bb.CG), D(bb.CGpart), D(bb.wf), rep_map_dev)) return rc;
  SolveBatch b;
  std::memset(&b, 0, sizeof(b));
  b.M = m->dv;
  b.G = D(bb.G); b.g_stride = (int64_t)h.n_tiles * TILE;
  b.colsum = D(bb.colsum); b.cs_stride = h.Ppad;
  b.mu = d->mu; b.N = (double)d->N; b.scheme = scheme; b.tol = tol; b.max_iter = max_iter;
  b.ws = D(bb.ws);
  b.iters = (int*)(base + bb.iters); b.status = (int*)(base + bb.status);
  b.cross = D(bb.CG); b.cross_stride = (int64_t)h.n_cross * TILE;
  b.phase = 2; b.rep_map = rep_map_dev;
  b.state = D(bb.state); b.state_stride = (int64_t)plspm::solver_state_doubles(m->dv); b.resume = 1;
  b.out_rows = out_rows; b.out_stride = h.n_out();
  const size_t smem = h.solver_core_smem_doubles() * sizeof(double);
  d->timer.begin(ST_SOLVE, st);
  solve_kernel<<<(unsigned)n_list, SOLVE_THREADS, smem, st>>>(b);
  d->timer.end(st);
  CK(cudaGetLastError());
  return 0;
}

// First and second moments of every replicate of a batch: Gram tiles + column sums.
static int launch_moments(plspm_data* d, int64_t nb, const uint32_t* counts_dev, const BatchBuffers& bb,
                          const BatchPlan& bp) {
  const HostModel& h = d->model->h;
  cudaStream_t st = d->stream;
  char* base = (char*)d->ws.ptr + d->ws_off;
  auto D = [&](size_t o) { return (double*)(base + o); };
  const bool i8 = d->i8_colsum && counts_dev;
  int8_t* c8 = (int8_t*)(base + bb.c8);
  if (i8 && !d->gram_mma) {
    d->timer.begin(ST_COLSUM, st);
    counts8_kernel<<<d->sm_count * 8, 256, 0, st>>>(counts_dev, d->N, d->Npad, nb, c8, (int*)(base + bb.ovf));
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  if (i8 && d->gram_mma && !d->img_ready) {
    d->timer.begin(ST_COLSUM, st);
    counts8_image_kernel<<<d->sm_count * 16, 256, 0, st>>>(counts_dev, d->N, nb, (int)((nb + 511) / 512), (d->N + 127) / 128,
                                                           (uint8_t*)(base + bb.c8img), (int*)(base + bb.ovf));
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  if (i8 && d->gram_mma) {
    // tcgen05 integer Gram: Gram tiles AND column sums of the batch in one kernel + the fp64 recombination
    if (int rc = gm_build_tables(d)) return rc;
    GramMmaParams gp;
    gp.mtiles = d->gm_tiles; gp.part = (longlong2*)(base + bb.gm_part);
    gp.nb = nb; gp.nb_pad = bp.gm_nb_pad; gp.N = d->N;
    gp.n_mtiles = d->gm_n_tiles; gp.n_groups = (int)((nb + 511) / 512); gp.ksplit = bp.gm_ksplit; gp.rows_per_cta = bp.gm_rows;
    gp.XsT = d->XsT; gp.ldx = d->ldx;
    gp.c8img = d->c8img_ovr ? d->c8img_ovr : (const uint8_t*)(base + bb.c8img);
    static const bool gstats = getenv("PLSPM_KERNEL_STATS") != nullptr;
    static unsigned long long* gstats_dev = nullptr;
    if (gstats && !gstats_dev) CK(cudaMalloc((void**)&gstats_dev, 16 * 8));
    if (gstats) CK(cudaMemsetAsync(gstats_dev, 0, 16 * 8, st));
    gp.stats = gstats ? gstats_dev : nullptr;
    const int64_t g_stride = (int64_t)h.n_tiles * TILE;
    CK(cudaMemsetAsync(D(bb.G), 0, (size_t)nb * g_stride * 8, st));
    CK(cudaMemsetAsync(D(bb.colsum), 0, (size_t)nb * h.Ppad * 8, st));
    const int64_t grid = (int64_t)gp.n_mtiles * gp.n_groups * gp.ksplit;
    if (grid > 0x7fffffff) return fail(PLSPM_ERR_UNSUPPORTED, "batch too large for one launch");
    CK(ensure_smem(gram_mma_kernel, (size_t)(gm_smem_bytes())));
    d->timer.begin(ST_GRAM_I8, st);
    gram_mma_kernel<<<(unsigned)grid, GM_THREADS, gm_smem_bytes(), st>>>(gp);
    d->timer.end(st);
    CK(cudaGetLastError());
    if (gstats) {
      unsigned long long hs[16];
      CK(cudaStreamSynchronize(st));
      CK(cudaMemcpy(hs, gstats_dev, sizeof(hs), cudaMemcpyDeviceToHost));
      const double ns = (double)std::max<unsigned long long>(hs[9], 1), nc = (double)grid;
      fprintf(stderr, "[gram_mma] grid %lld, stages/CTA %.0f | per stage: issuer loop %.0f clk (waits: c8 tile %.0f, digits %.0f); "
              "producer waits %.0f; generator loop %.0f (operand wait %.0f + digits %.0f, buffer wait %.0f) | epilogue %.0f clk per CTA\n",
              (long long)grid, ns / nc, hs[0] / ns, hs[2] / ns, hs[3] / ns, hs[1] / ns, hs[6] / ns, hs[8] / ns, (hs[5] - hs[8]) / ns,
              hs[4] / ns, hs[7] / nc);
    }
    d->timer.begin(ST_FINALIZE, st);
    gram_finalize_kernel<<<d->sm_count * 8, 256, 0, st>>>(gp.part, nb, gp.nb_pad, gp.n_mtiles * GM_PAIRS_PER_TILE, gp.ksplit,
                                                          (const GramOut*)d->gm_outs, d->gm_n_outs, d->xunit, (double)d->N,
                                                          g_stride, D(bb.G), h.Ppad, D(bb.colsum));
    d->timer.end(st);
    CK(cudaGetLastError());
    return 0;
  }
  bool gram_done = false;
  if (i8 && d->n_zcols && d->z_model == d->model) {
    // all Gram tiles of the batch: counts8 [nb x N] x pair-product planes [N x 6 n_zcols], exact int32 sums
    int32_t* zs = (int32_t*)(base + bb.zs32);
    const int32_t one = 1, zero = 0;
    const int gemm_m = I8_DIGITS * d->n_zcols;
    const int64_t g_stride = (int64_t)h.n_tiles * TILE;
    const int64_t step = d->Z8 ? I8_KCHUNK : d->z_chunk_rows;
    CK(cudaMemsetAsync(D(bb.G), 0, (size_t)nb * g_stride * 8, st));
    gram_done = true;
    for (int64_t k0 = 0; k0 < d->Npad; k0 += step) {
      const int kc = (int)std::min<int64_t>(step, d->Npad - k0);
      const int8_t* planes = d->Z8 ? d->Z8 + k0 : (const int8_t*)(base + bb.zchunk);
      const int64_t ld = d->Z8 ? d->Npad : kc;
      if (!d->Z8) {
        d->timer.begin(ST_GRAM_I8, st);
        zdigits_kernel<<<dim3((d->n_zcols + 31) / 32, (unsigned)((kc + 127) / 128)), 256, 0, st>>>(
            d->X, d->N, h.Ppad, k0, kc, d->n_zcols, d->zp, d->zq, d->zqscale, (int8_t*)(base + bb.zchunk));
        d->timer.end(st);
        CK(cudaGetLastError());
      }
      d->timer.begin(ST_GRAM_I8, st);
      cublasStatus_t cs = cublasGemmEx(d->blas, CUBLAS_OP_T, CUBLAS_OP_N, gemm_m, (int)nb, kc, &one, planes, CUDA_R_8I,
                                       (int)ld, c8 + k0, CUDA_R_8I, (int)d->Npad, &zero, zs, CUDA_R_32I, gemm_m,
                                       CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
      d->timer.end(st);
      if (cs != CUBLAS_STATUS_SUCCESS) { gram_done = false; break; }
      d->timer.begin(ST_GRAM_I8, st);
      zcombine_kernel<<<d->sm_count * 8, 256, 0, st>>>(zs, nb, d->n_zcols, d->zdscale, d->zdst, d->zdst2, k0 > 0 ? 1 : 0,
                                                      g_stride, D(bb.G));
      d->timer.end(st);
      CK(cudaGetLastError());
    }
    if (!gram_done) d->z_model = nullptr;  // no int8 GEMM for this shape: fp64 kernel from now on
  }
  if (!gram_done)
    if (int rc = launch_stream(d, false, nb, counts_dev, bp.gram, D(bb.G), D(bb.Gpart), nullptr)) return rc;
  if (i8) {
    int32_t* s32 = (int32_t*)(base + bb.s32);
    const int32_t one = 1, zero = 0;
    const int gemm_m = I8_DIGITS * h.Ppad;
    bool ok = true;
    for (int64_t k0 = 0; k0 < d->Npad && ok; k0 += I8_KCHUNK) {
      const int kc = (int)std::min<int64_t>(I8_KCHUNK, d->Npad - k0);
      d->timer.begin(ST_COLSUM, st);
      cublasStatus_t cs = cublasGemmEx(d->blas, CUBLAS_OP_T, CUBLAS_OP_N, gemm_m, (int)nb, kc, &one, d->D8 + k0, CUDA_R_8I,
                                       (int)d->Npad, c8 + k0, CUDA_R_8I, (int)d->Npad, &zero, s32, CUDA_R_32I, gemm_m,
                                       CUBLAS_COMPUTE_32I, CUBLAS_GEMM_DEFAULT);
      d->timer.end(st);
      if (cs != CUBLAS_STATUS_SUCCESS) { ok = false; break; }
      d->timer.begin(ST_COLSUM, st);
      digits_combine_kernel<<<d->sm_count * 2, 256, 0, st>>>(s32, nb, h.Ppad, d->dscale, k0 > 0 ? 1 : 0, D(bb.colsum));
      d->timer.end(st);
      CK(cudaGetLastError());
    }
    if (ok) return 0;
    d->i8_colsum = false;  // this cuBLAS build has no int8 GEMM for the shape: fp64 kernel from now on
  }
  if (!counts_dev && nb == 1 && d->colsum0) {  // unweighted fit: the sums of the centred columns are a by-product of the upload
    CK(cudaMemcpyAsync(D(bb.colsum), d->colsum0, (size_t)h.Ppad * 8, cudaMemcpyDeviceToDevice, st));
    return 0;
  }
  dim3 grid_cs((h.Ppad + CS_COLS - 1) / CS_COLS, (unsigned)((nb + CS_REPS - 1) / CS_REPS), bp.cs_chunks);
  d->timer.begin(ST_COLSUM, st);
  const size_t cs_smem_bytes = (size_t)(CS_ROWS * CS_COLS + CS_ROWS * CS_REPS) * 8;
  CK(ensure_smem(colsum_kernel, (size_t)(cs_smem_bytes)));
  colsum_kernel<<<grid_cs, 256, cs_smem_bytes, st>>>(d->X, counts_dev, d->N, h.Ppad, nb, bp.cs_chunks, bp.cs_chunk_rows,
                                                     bp.cs_chunks > 1 ? D(bb.cspart) : D(bb.colsum));
  d->timer.end(st);
  if (bp.cs_chunks > 1) {
    d->timer.begin(ST_REDUCE, st);
    reduce_chunks_kernel<<<d->sm_count, 256, 0, st>>>(D(bb.cspart), nb, bp.cs_chunks, h.Ppad, D(bb.colsum));
    d->timer.end(st);
  }
  CK(cudaGetLastError());
  return 0;
}

// Numeric non-metric path (solver_num.h): the stopping rule is evaluated on the scores, so the outer
// iteration is driven from here -- one num_step launch (all unfinished replicates advance by one
// iteration, or finish) and one criterion pass over X per iteration.
static int run_batch_num(plspm_data* d, int64_t nb, const uint32_t* counts_dev, const BatchBuffers& bb, int scheme,
                         double tol, int max_iter, const BatchPlan& bp, double* out_rows, bool single_fit) {
  const plspm_model* m = d->model;
  const HostModel& h = m->h;
  cudaStream_t st = d->stream;
  char* base = (char*)d->ws.ptr + d->ws_off;
  auto D = [&](size_t o) { return (double*)(base + o); };
  if (int rc = launch_moments(d, nb, counts_dev, bb, bp)) return rc;
  NumBatch b;
  std::memset(&b, 0, sizeof(b));
  b.M = m->dv;
  b.G = D(bb.G); b.g_stride = (int64_t)h.n_tiles * TILE;
  b.colsum = D(bb.colsum);
  b.N = (double)d->N; b.scheme = scheme; b.tol = tol; b.max_iter = max_iter;
  b.conv_part = D(bb.num_cpart); b.n_conv_part = (int)bp.cv_gx;
  b.conv_main = D(bb.num_cmain);
  b.ws = D(bb.ws);
  b.a = D(bb.num_a); b.coef_old = D(bb.num_co); b.coef_new = D(bb.num_cn);
  b.shift_old = D(bb.num_so); b.shift_new = D(bb.num_sn);
  b.meta = (int*)(base + bb.num_meta); b.n_done = (int*)(base + bb.num_done);
  b.out_rows = out_rows; b.out_stride = h.n_out();
  b.iters = (int*)(base + bb.iters); b.status = (int*)(base + bb.status);
  if (single_fit) {
    b.weights = D(bb.weights); b.loadings = D(bb.loadings); b.r2 = D(bb.r2); b.paths = D(bb.paths);
    b.total = D(bb.totalfx); b.crossloadings = D(bb.crossl); b.score_coef = D(bb.coef); b.score_shift = D(bb.shift);
  }
  CK(cudaMemsetAsync(b.meta, 0, (size_t)nb * 16, st));
  CK(cudaMemsetAsync(b.n_done, 0, 8, st));
  CK(cudaMemsetAsync(D(bb.num_cpart), 0, (size_t)nb * bp.cv_gx * 8, st));
  CK(cudaMemsetAsync(D(bb.num_cmain), 0, (size_t)nb * 8, st));
  const size_t smem = h.solver_smem_doubles() * sizeof(double);
  if (smem > (size_t)d->max_smem) return fail(PLSPM_ERR_UNSUPPORTED, "model too large for the solver's shared memory");
  CK(ensure_smem(num_step_kernel, (size_t)(smem)));
  CK(ensure_smem(conv_kernel<float>, (size_t)(bp.cv_smem)));
  CK(ensure_smem(conv_kernel<double>, (size_t)(bp.cv_smem)));
  const unsigned gy = (unsigned)((nb + bp.cv_reps_per_cta - 1) / bp.cv_reps_per_cta);
  int done = 0;
  for (int step = 0; step < max_iter + 4; ++step) {
    d->timer.begin(ST_SOLVE, st);
    num_step_kernel<<<(unsigned)nb, 128, smem, st>>>(b);
    d->timer.end(st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&done, b.n_done, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (done >= nb) break;
    d->timer.begin(ST_CONV, st);
    if (bp.cv_f32)
      conv_kernel<float><<<dim3(bp.cv_gx, gy), SG_THREADS, bp.cv_smem, st>>>(
          d->Xf, d->X, counts_dev, b.coef_old, b.coef_new, b.shift_old, b.shift_new, b.meta, d->N, h.Ppad, h.L,
          m->dv.lv_off, m->dv.lv_k, bp.cv_nsl_pad, bp.cv_rows, nb, D(bb.num_cpart));
    else
      conv_kernel<double><<<dim3(bp.cv_gx, gy), SG_THREADS, bp.cv_smem, st>>>(
          d->X, d->X, counts_dev, b.coef_old, b.coef_new, b.shift_old, b.shift_new, b.meta, d->N, h.Ppad, h.L,
          m->dv.lv_off, m->dv.lv_k, bp.cv_nsl_pad, bp.cv_rows, nb, D(bb.num_cpart));
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  if (done < nb) return fail(PLSPM_ERR_CUDA, "numeric non-metric iteration did not terminate");
  return 0;
}

// Shared implementation of fit (counts == null, one "replicate") and bootstrap batches.
static int run_batch(plspm_data* d, int64_t nb, const uint32_t* counts_dev, const BatchBuffers& bb, int scheme,
                     double tol, int max_iter, const BatchPlan& bp, double* out_rows, bool single_fit) {
  const plspm_model* m = d->model;
  const HostModel& h = m->h;
  const bool want_fast = !h.full && d->fast_vote && !single_fit && (int64_t)(nb + 8) * h.L < (1 << 30);
  // fused tcgen05 vote: needs the int8 multiplicities of the batch (built by launch_moments)
  const bool use_mma = want_fast && d->mma_vote && d->i8_colsum && counts_dev != nullptr;
  const bool use_fast = use_mma || (want_fast && d->Xf && d->blas);
  cudaStream_t st = d->stream;
  char* base = (char*)d->ws.ptr + d->ws_off;
  auto D = [&](size_t o) { return (double*)(base + o); };
  if (int rc = launch_moments(d, nb, counts_dev, bb, bp)) return rc;
  SolveBatch b;
  std::memset(&b, 0, sizeof(b));
  if (d->imp && !single_fit) {
    // missing values: moments of the augmented matrix -> moments of the per-replicate imputed data -> base-model solve
    ImputeCtx& ic = *d->imp;
    const HostModel& hb = ic.base->h;
    if (ic.cap < nb) {
      for (void* q : {(void*)ic.G, (void*)ic.colsum, (void*)ic.ws})
        if (q) g_pool.release(q);
      ic.G = ic.colsum = ic.ws = nullptr;
      ic.cap = 0;
      CK(g_pool.alloc((void**)&ic.G, (size_t)nb * hb.n_tiles * TILE * 8));
      CK(g_pool.alloc((void**)&ic.colsum, (size_t)nb * hb.Ppad * 8));
      CK(g_pool.alloc((void**)&ic.ws, (size_t)nb * std::max(hb.ws_doubles, 1) * 8));
      ic.cap = nb;
    }
    d->timer.begin(ST_FINALIZE, st);
    impute_moments_kernel<<<(unsigned)nb, 256, (size_t)4 * hb.Ppad * 8, st>>>(
        m->dv, ic.base->dv, D(bb.G), (int64_t)h.n_tiles * TILE, D(bb.colsum), h.Ppad, d->mu, ic.ax, ic.am, (double)d->N, ic.G,
        (int64_t)hb.n_tiles * TILE, ic.colsum, hb.Ppad);
    d->timer.end(st);
    CK(cudaGetLastError());
    b.M = ic.base->dv;
    b.G = ic.G; b.g_stride = (int64_t)hb.n_tiles * TILE;
    b.colsum = ic.colsum; b.cs_stride = hb.Ppad;
    b.mu = ic.mu_base; b.N = (double)d->N; b.scheme = scheme; b.tol = tol; b.max_iter = max_iter;
    b.ws = ic.ws;
    b.iters = (int*)(base + bb.iters); b.status = (int*)(base + bb.status);
    b.phase = 0;
    b.out_rows = out_rows; b.out_stride = hb.n_out();
    const size_t smem_b = hb.solver_core_smem_doubles() * sizeof(double);
    if (smem_b > (size_t)d->max_smem) return fail(PLSPM_ERR_UNSUPPORTED, "model too large for the solver's shared memory");
    CK(ensure_smem(solve_kernel, (size_t)(smem_b)));
    d->timer.begin(ST_SOLVE, st);
    solve_kernel<<<(unsigned)nb, SOLVE_THREADS, smem_b, st>>>(b);
    d->timer.end(st);
    CK(cudaGetLastError());
    return 0;
  }
  b.M = m->dv;
  b.G = D(bb.G); b.g_stride = (int64_t)h.n_tiles * TILE;
  b.colsum = D(bb.colsum); b.cs_stride = h.Ppad;
  b.mu = d->mu; b.N = (double)d->N; b.scheme = scheme; b.tol = tol; b.max_iter = max_iter;
  b.ws = D(bb.ws);
  b.iters = (int*)(base + bb.iters); b.status = (int*)(base + bb.status);
  if (!h.full) { b.state = D(bb.state); b.state_stride = (int64_t)plspm::solver_state_doubles(m->dv); }
  b.wf = h.full ? nullptr : D(bb.wf);
  b.cross = h.full ? nullptr : D(bb.CG); b.cross_stride = (int64_t)h.n_cross * TILE;
  const size_t smem = h.solver_core_smem_doubles() * sizeof(double);
  if (smem > (size_t)d->max_smem) return fail(PLSPM_ERR_UNSUPPORTED, "model too large for the solver's shared memory");
  CK(ensure_smem(solve_kernel, (size_t)(smem)));
  if (!h.full) {
    // sparse tile set: final weights first, then the P x L cross-moment pass for the sign vote
    b.phase = 1;
    b.sh = D(bb.sh);
    d->timer.begin(ST_SOLVE, st);
    solve_kernel<<<(unsigned)nb, SOLVE_THREADS, smem, st>>>(b);
    d->timer.end(st);
    CK(cudaGetLastError());
    b.sh = nullptr;
    if (use_mma) {
      // fused tensor-core sign vote (kernels_vote_mma.cuh): scores, multiplicity scaling and the P x L contraction
      // in one kernel; the partial accumulators of the row ranges are added into Cf
      float* Cf = (float*)(base + bb.Cf);
      const int64_t ldl = (nb + 7) / 8 * 8;
      VoteMmaParams vp;
      vp.wf = D(bb.wf); vp.inv_sd = d->inv_sd; vp.lv_off = m->dv.lv_off; vp.lv_k = m->dv.lv_k; vp.lv_blk = d->lv_blk; vp.Cf = Cf;
      vp.xt_img = d->xt_img; vp.xl_img = d->xl_img; vp.c8_img = d->c8vote_ovr ? d->c8vote_ovr : (const uint8_t*)(base + bb.c8vote);
      vp.n_blocks = d->vm_n_blocks; vp.n_rep_tiles_img = (int)((nb + 127) / 128);
      if (!d->img_ready) {
        d->timer.begin(ST_COLSUM, st);
        vote_c8_image_kernel<<<d->sm_count * 16, 256, 0, st>>>(counts_dev, d->N, nb, vp.n_rep_tiles_img, d->n_chunks64,
                                                               (uint8_t*)(base + bb.c8vote));
        d->timer.end(st);
        CK(cudaGetLastError());
      }
      vp.nb = nb; vp.ldl = ldl; vp.N = d->N; vp.L = h.L; vp.Ppad = h.Ppad;
      vp.n_pchunks = (h.Ppad + 255) / 256;
      vp.k16_max = std::max(1, (h.kmax + 15) / 16);
      vp.n_rep_tiles = (int)((nb + 127) / 128);
      // row ranges: at most VM_MAX_ROWS rows per accumulator (truncating fp32 accumulation); among the admissible
      // splits take the one whose CTA count fills whole waves best
      const int64_t tiles0 = (int64_t)vp.n_rep_tiles * h.L * vp.n_pchunks;
      const int kmin = (int)((d->N + VM_MAX_ROWS - 1) / VM_MAX_ROWS);
      double best = 1e300;
      for (int k = kmin; k <= kmin + 24; ++k) {
        const int64_t rows = ((d->N + k - 1) / k + VM_STAGE_ROWS - 1) / VM_STAGE_ROWS * VM_STAGE_ROWS;
        if (rows > VM_MAX_ROWS || ((int64_t)k * rows - d->N >= rows && k > kmin)) continue;  // too long / an empty last range
        const double cost = (double)((tiles0 * k + d->sm_count - 1) / d->sm_count) * (double)(rows + 768);
        if (cost < best) { best = cost; vp.ksplit = k; vp.rows_per_cta = (int)rows; }
      }
      // PLSPM_KERNEL_STATS=1: cycles every role of the kernel spent waiting on each barrier (printed per launch)
      static const bool kstats = getenv("PLSPM_KERNEL_STATS") != nullptr;
      static unsigned long long* kstats_dev = nullptr;
      if (kstats && !kstats_dev) CK(cudaMalloc((void**)&kstats_dev, 16 * 8));
      if (kstats) CK(cudaMemsetAsync(kstats_dev, 0, 16 * 8, st));
      vp.stats = kstats ? kstats_dev : nullptr;
      vp.stages = vm_stages(vp.k16_max, d->max_smem);
      const size_t vm_smem = vm_smem_bytes(vp.k16_max, vp.stages);
      CK(ensure_smem(vote_mma_kernel, (size_t)(vm_smem)));
      CK(cudaMemsetAsync(Cf, 0, (size_t)ldl * h.L * h.Ppad * sizeof(float), st));
      const int64_t grid = tiles0 * vp.ksplit;
      if (grid > 0x7fffffff) return fail(PLSPM_ERR_UNSUPPORTED, "batch too large for one launch");
      d->timer.begin(ST_CROSS, st);
      vote_mma_kernel<<<(unsigned)grid, VM_THREADS, vm_smem, st>>>(vp);
      d->timer.end(st);
      CK(cudaGetLastError());
      if (kstats) {
        unsigned long long hs[16];
        CK(cudaStreamSynchronize(st));
        CK(cudaMemcpy(hs, kstats_dev, sizeof(hs), cudaMemcpyDeviceToHost));
        const double nc = (double)std::max<unsigned long long>(hs[9], 1);
        fprintf(stderr, "[vote_mma] grid %lld, chunks/CTA %.0f | per chunk: score issuer %.0f clk, vote issuer %.0f clk; waits: "
                "producer(empty) %.0f, vote issuer a_full %.0f, score issuer full %.0f d1_empty %.0f, epilogue (per own chunk) "
                "d1_full %.0f a_free %.0f\n", (long long)grid, nc / (double)grid, hs[0] / nc, hs[8] / nc, hs[1] / nc, hs[2] / nc,
                hs[3] / nc, hs[4] / nc, 2.0 * hs[5] / nc, 2.0 * hs[7] / nc);
      }
      b.fast_cross = Cf; b.inv_sd = d->inv_sd; b.fast_nb = ldl; b.fast_uncentred = 1;
    } else if (use_fast) {
      // legacy tensor-core sign vote: E[p][b][l] = sum_i xh_ip * fp16(c_bi t_bil), fp32 accumulate, in chunks
      // of FAST_RC rows (cuBLAS: a plain fp16 GEMM, m = nb*L, n = Ppad, k = rows of the chunk)
      __half* BT = (__half*)(base + bb.BT);
      float* Cf = (float*)(base + bb.Cf);
      int nsl_pad = 1;
      while (nsl_pad * SLOT < h.kmax) nsl_pad <<= 1;
      // replicate lanes per CTA: SG_THREADS / (L * nsl_pad); with single-slot blocks the warp layout is
      // (8 LVs x 4 lanes) x ceil(L/8) LV groups, i.e. lanes come in fours
      int nbl_host = SG_THREADS / (h.L * nsl_pad);
      if (nsl_pad == 1 && nbl_host >= 4) nbl_host = (SG_THREADS / 32 / ((h.L + 7) / 8)) * 4;
      const int reps_per_cta = std::max(1, nbl_host) * SG_RPT;
      // rows per staged tile: at most 64, and small enough for two CTAs per SM
      const size_t sg_row_bytes = (size_t)h.Ppad * 4 + (size_t)reps_per_cta * 4;
      const int SG_ROWS = (int)std::max<size_t>(1, std::min<size_t>(SG_MAX_ROWS, (size_t)(d->max_smem / 2 - 8192) / sg_row_bytes));
      const size_t sg_smem = (size_t)SG_ROWS * sg_row_bytes;
      auto sg_kernel = (nsl_pad == 1) ? scoregen_kernel<true> : scoregen_kernel<false>;
      CK(ensure_smem(sg_kernel, (size_t)(sg_smem)));
      const float one = 1.f, zero = 0.f;
      const int64_t ldl = (nb + 7) / 8 * 8;
      const int gemm_m = (int)(ldl * h.L);
      // padded replicate columns of the score buffer are only written up to the CTA grid's reach: clear once
      CK(cudaMemsetAsync(BT, 0, (size_t)gemm_m * FAST_RC * sizeof(__half), st));
      for (int64_t i0 = 0; i0 < d->N; i0 += FAST_RC) {
        const int rc = (int)std::min<int64_t>(FAST_RC, d->N - i0);
        const unsigned gy = (unsigned)((nb + reps_per_cta - 1) / reps_per_cta);
        const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((rc + SG_ROWS - 1) / SG_ROWS,
                                                                            (4 * d->sm_count + gy - 1) / gy));
        dim3 grid_sg(gx, gy);
        d->timer.begin(ST_SCOREGEN, st);
        sg_kernel<<<grid_sg, SG_THREADS, sg_smem, st>>>(d->Xf, counts_dev, D(bb.wf), D(bb.sh), d->N, h.Ppad, h.L,
                                                             m->dv.lv_off, m->dv.lv_k, nsl_pad, SG_ROWS, nb, ldl, i0, rc,
                                                             BT);
        d->timer.end(st);
        CK(cudaGetLastError());
        // C[nb*L x Ppad] (+)= B^T-free NT product: A = scores stored [nb*L x rc] (column-major view of the
        // row-major [rc][nb*L] buffer), B = xh chunk stored [Ppad x rc]
        d->timer.begin(ST_CROSS, st);
        cublasStatus_t cs = cublasGemmEx(d->blas, CUBLAS_OP_N, CUBLAS_OP_T, gemm_m, h.Ppad, rc, &one, BT, CUDA_R_16F,
                                         gemm_m, d->Xh + i0 * h.Ppad, CUDA_R_16F, h.Ppad, i0 == 0 ? &zero : &one, Cf,
                                         CUDA_R_32F, gemm_m, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT);
        d->timer.end(st);
        if (cs != CUBLAS_STATUS_SUCCESS) return fail(PLSPM_ERR_CUDA, "cublasGemmEx failed: " + std::to_string((int)cs));
      }
      b.fast_cross = Cf; b.inv_sd = d->inv_sd; b.fast_nb = ldl;
    } else {
      if (int rc = launch_stream(d, true, nb, counts_dev, bp.cross, D(bb.CG), D(bb.CGpart), D(bb.wf))) return rc;
    }
  }
  b.phase = h.full ? 0 : (use_fast ? 3 : 2);
  b.resume = h.full ? 0 : 1;  // phase 1 of this batch left the converged state in bb.state
  b.out_rows = out_rows; b.out_stride = h.n_out();
  if (single_fit) {
    b.weights = D(bb.weights); b.loadings = D(bb.loadings); b.r2 = D(bb.r2); b.paths = D(bb.paths);
    b.total = D(bb.totalfx); b.crossloadings = D(bb.crossl); b.score_coef = D(bb.coef); b.score_shift = D(bb.shift);
  }
  d->timer.begin(ST_SOLVE, st);
  solve_kernel<<<(unsigned)nb, SOLVE_THREADS, smem, st>>>(b);
  d->timer.end(st);
  CK(cudaGetLastError());
  return 0;
}

// Small page-locked staging buffers for the per-batch read-back (status, iterations, overflow flag): cached
// process-wide, because plspm_bootstrap_host creates a handle per call and cudaMallocHost costs ~0.1 ms.
struct PinnedStage {
  int* ints = nullptr;
  size_t bytes = 0;
  static std::mutex& mu() { static std::mutex m; return m; }
  static std::vector<std::pair<void*, size_t>>& cache() { static std::vector<std::pair<void*, size_t>> c; return c; }
  explicit PinnedStage(size_t need) {
    need = std::max<size_t>((need + 65535) / 65536 * 65536, 65536);
    {
      std::lock_guard<std::mutex> lk(mu());
      auto& c = cache();
      for (size_t i = 0; i < c.size(); ++i)
        if (c[i].second >= need) {
          ints = (int*)c[i].first; bytes = c[i].second;
          c.erase(c.begin() + i);
          return;
        }
    }
    if (cudaMallocHost((void**)&ints, need) == cudaSuccess) bytes = need;
    else ints = nullptr;
  }
  ~PinnedStage() {
    if (!ints) return;
    std::lock_guard<std::mutex> lk(mu());
    if (cache().size() < 16) cache().push_back({(void*)ints, bytes});
    else cudaFreeHost(ints);
  }
  PinnedStage(const PinnedStage&) = delete;
  PinnedStage& operator=(const PinnedStage&) = delete;
};

int plspm_fit(const plspm_model* m, plspm_data* d, int32_t scheme, double tol, int32_t max_iter, double* weights,
              double* loadings, double* r_squared, double* paths, double* total_effects, double* crossloadings,
              double* scores, int32_t* iters, int32_t* status) {
  if (!m || !d || !same_layout(d->model, m)) return fail(PLSPM_ERR_INVALID, "model/data mismatch");
  if (scheme < 0 || scheme > 2) return fail(PLSPM_ERR_INVALID, "unknown scheme");
  d->model = m;  // the handle carries the model of the call in flight (one call at a time per handle)
  const HostModel& h = m->h;
  const int64_t N = d->N;
  const size_t L = h.L, P = h.P;
  BatchPlan bp;
  if (int rc = plan_batch(d, 1, bp)) return rc;
  const BatchBuffers bb = layout_batch(d, 1, bp, false, false, false, true, scores != nullptr);
  if (int rc = ws_reserve(d, bb.total)) return rc;
  char* base = (char*)d->ws.ptr + d->ws_off;
  auto D = [&](size_t o) { return (double*)(base + o); };
  cudaStream_t st = d->stream;
  if (m->numeric) {
    if (crossloadings && !h.full)
      return fail(PLSPM_ERR_UNSUPPORTED, "numeric non-metric fit: crossloadings need a model with the full tile set");
    if (int rc = run_batch_num(d, 1, nullptr, bb, scheme, tol, max_iter, bp, nullptr, true)) return rc;
  } else if (int rc = run_batch(d, 1, nullptr, bb, scheme, tol, max_iter, bp, nullptr, true)) {
    return rc;
  }
  if (scores) {
    d->timer.begin(ST_SCORES, st);
    scores_kernel<<<d->sm_count * 8, 256, 0, st>>>(d->X, N, h.Ppad, h.L, m->dv.lv_off, m->dv.lv_k, D(bb.coef),
                                                   D(bb.shift), D(bb.scores));
    d->timer.end(st);
    CK(cudaGetLastError());
  }
  // The small outputs are consecutive in the workspace (layout_batch): ONE copy into page-locked staging instead of
  // eight copies into the caller's pageable arrays (each of those is a synchronous round trip of ~10 us, which was
  // most of the host-side gap of a 0.8 ms fit).  iters / status sit next to each other as well.
  const size_t small_bytes = bb.crossl + P * L * 8 - bb.weights, tail_bytes = bb.status + 4 - bb.iters;
  PinnedStage stage(small_bytes + tail_bytes);
  if (!stage.ints) return fail(PLSPM_ERR_NOMEM, "pinned staging buffer");
  char* hs = (char*)stage.ints;
  CK(cudaMemcpyAsync(hs, base + bb.weights, small_bytes, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(hs + small_bytes, base + bb.iters, tail_bytes, cudaMemcpyDeviceToHost, st));
  if (scores) CK(cudaMemcpyAsync(scores, D(bb.scores), (size_t)N * L * 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  d->timer.collect();
  auto from = [&](size_t off) { return hs + (off - bb.weights); };
  if (weights) std::memcpy(weights, from(bb.weights), P * 8);
  if (loadings) std::memcpy(loadings, from(bb.loadings), P * 8);
  if (r_squared) std::memcpy(r_squared, from(bb.r2), L * 8);
  if (paths) std::memcpy(paths, from(bb.paths), L * L * 8);
  if (total_effects) std::memcpy(total_effects, from(bb.totalfx), L * L * 8);
  if (crossloadings) std::memcpy(crossloadings, from(bb.crossl), P * L * 8);
  if (iters) std::memcpy(iters, hs + small_bytes, 4);
  if (status) std::memcpy(status, hs + small_bytes + (bb.status - bb.iters), 4);
  return PLSPM_OK;
}

// Replicates per batch of plspm_bootstrap: bounds the workspace and keeps the grids whole waves.  A function of the
// model, the row count and the route flags only, so plspm_bootstrap_host can know the first batch before the data
// handle exists (prefetch of the multiplicity images, below).
static int64_t bootstrap_batch_size(const plspm_model* m, int64_t N, size_t n_out, bool with_idx, bool gram_mma, bool i8_colsum,
                                    int64_t Npad, int n_zcols, bool z_resident, int sm_count, int64_t rep_count) {
  const HostModel& h = m->h;
  const bool vote = !h.full && !m->numeric;
  const size_t per_rep = (size_t)N * 4 + ((size_t)h.n_tiles * TILE + (vote ? (size_t)h.n_cross * TILE : 0) +
                                          (m->numeric ? 6 : 2) * h.Ppad + h.ws_doubles + n_out) * 8 +
                         (with_idx ? (size_t)N * 4 : 0) +
                         (gram_mma ? (size_t)Npad + (size_t)gm_count_tiles(h) * GM_PAIRS_PER_TILE * sizeof(longlong2) * 3 +
                                         (vote ? (size_t)h.L * h.Ppad * 4 : 0)
                                   : i8_colsum ? (size_t)Npad + I8_DIGITS * ((size_t)h.Ppad + n_zcols) * 4 : 0) + 64;
  // the planes of a chunk are regenerated for every batch in streaming mode: large batches amortise that
  const size_t ws_budget = n_zcols && !z_resident ? (size_t)12 << 30 : gram_mma ? (size_t)3 << 30 : (size_t)1536 << 20;
  int64_t nb_max = std::max<int64_t>(1, (int64_t)(ws_budget / per_rep));
  if (getenv("PLSPM_MAX_BATCH")) nb_max = std::max<int64_t>(1, std::min<int64_t>(nb_max, atoll(getenv("PLSPM_MAX_BATCH"))));
  nb_max = std::min<int64_t>(nb_max, rep_count);
  const int64_t wave = (int64_t)sm_count * GRAM_WARPS;  // warp items per wave
  const int64_t items_per_rep = vote ? std::max(h.n_tg, h.n_tg_cross) : h.n_tg;
  if (gram_mma) {
    // tensor-core kernels: CTAs own 512 (Gram) / 128 (sign vote) replicates and row ranges even out the waves.
    // Batches of equal size (no small tail batch), whole 128-replicate tiles where the run is large enough.
    if (nb_max >= 512) nb_max = nb_max / 512 * 512;
    const int64_t n_batches = (rep_count + nb_max - 1) / nb_max;
    int64_t even = (rep_count + n_batches - 1) / n_batches;
    if (even >= 128) even = (even + 127) / 128 * 128;
    nb_max = std::min(nb_max, even);
  } else if (nb_max * items_per_rep > wave) {
    int64_t waves = nb_max * items_per_rep / wave;
    nb_max = std::max<int64_t>(1, waves * wave / items_per_rep);
  }
  return nb_max;
}

// The multiplicity images of the FIRST batch of a plspm_bootstrap_host call do not depend on the data: they are
// generated on a side stream while the observation matrix crosses PCIe (3.8 of the 11.7 ms of a c3 call), into pooled
// buffers the first batch then reads in place of its workspace images.  Speculative: if the data handle ends up on
// another route (heavy-tailed data, small N) the images are simply dropped.
struct ImagePrefetch {
  uint8_t *c8img = nullptr, *c8vote = nullptr;
  int* ovf = nullptr;
  cudaEvent_t done = nullptr;
  int64_t nb = 0, rep_begin = 0;
  uint64_t seed = 0;
  bool valid = false;
};
static cudaStream_t side_stream() {
  static thread_local cudaStream_t s = nullptr;
  static thread_local int dev_of = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!s || dev_of != dev) {
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) s = nullptr;
    dev_of = dev;
  }
  return s;
}
static void prefetch_release(ImagePrefetch& pf) {
  if (pf.c8img) g_pool.release(pf.c8img);
  if (pf.c8vote) g_pool.release(pf.c8vote);
  if (pf.ovf) g_pool.release(pf.ovf);
  if (pf.done) cudaEventDestroy(pf.done);
  pf = ImagePrefetch();
}
static void prefetch_images(const plspm_model* m, int64_t N, int64_t rep_begin, int64_t rep_count, uint64_t seed, ImagePrefetch& pf) {
  static const bool off = getenv("PLSPM_GRAM") || getenv("PLSPM_VOTE") || getenv("PLSPM_COUNTS") || getenv("PLSPM_PREFETCH_OFF");
  const HostModel& h = m->h;
  const bool vote = !h.full && !m->numeric;
  if (off || m->numeric || N < 4096 || N >= ((int64_t)1 << 31) - 256 || rep_count < 1 || (vote && h.kmax > 16 * VM_MAX_K16)) return;
  int dev = 0, sm_count = 148;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return;
  cudaStream_t ss = side_stream();
  if (!ss) return;
  const int64_t nb = bootstrap_batch_size(m, N, h.n_out(), false, true, true, (N + 15) / 16 * 16, 0, false, sm_count, rep_count);
  const int n_groups = (int)((nb + 511) / 512), n_rep_tiles = vote ? (int)((nb + 127) / 128) : 0;
  const int64_t n_pad = (N + 127) / 128 * 128, n_chunks64 = (N + 63) / 64;
  const int64_t n_ranges = (n_pad + RI_MAX_ROWS - 1) / RI_MAX_ROWS;
  const int64_t range_rows = ((n_pad + n_ranges - 1) / n_ranges + 127) / 128 * 128;
  if (g_pool.alloc((void**)&pf.c8img, (size_t)(n_pad / 128) * n_groups * 65536) != cudaSuccess ||
      (vote && g_pool.alloc((void**)&pf.c8vote, (size_t)n_chunks64 * n_rep_tiles * 8192) != cudaSuccess) ||
      g_pool.alloc((void**)&pf.ovf, 8) != cudaSuccess || cudaEventCreateWithFlags(&pf.done, cudaEventDisableTiming) != cudaSuccess ||
      ensure_smem(resample_images_kernel, (size_t)range_rows) != cudaSuccess) {
    prefetch_release(pf);
    cudaGetLastError();
    return;
  }
  cudaMemsetAsync(pf.ovf, 0, 8, ss);
  resample_images_kernel<<<dim3((unsigned)n_groups * 512, (unsigned)n_ranges), RI_THREADS, (size_t)range_rows, ss>>>(
      nullptr, N, nb, rep_begin, seed, range_rows, n_groups, n_rep_tiles, n_chunks64, pf.c8img, pf.c8vote, pf.ovf);
  cudaEventRecord(pf.done, ss);
  if (cudaGetLastError() != cudaSuccess) { prefetch_release(pf); return; }
  pf.nb = nb; pf.rep_begin = rep_begin; pf.seed = seed; pf.valid = true;
}

int plspm_bootstrap(const plspm_model* m, plspm_data* d, int32_t scheme, double tol, int32_t max_iter,
                    int64_t rep_begin, int64_t rep_count, uint64_t seed, const int32_t* idx, double* out,
                    int32_t out_is_device, int32_t* status, int32_t* iters) {
  if (!m || !d || !same_layout(d->model, m)) return fail(PLSPM_ERR_INVALID, "model/data mismatch");
  if (scheme < 0 || scheme > 2) return fail(PLSPM_ERR_INVALID, "unknown scheme");
  if (rep_count < 0 || !out) return fail(PLSPM_ERR_INVALID, "bad replicate range / null output");
  if (rep_count == 0) return PLSPM_OK;
  d->model = m;
  const HostModel& h = m->h;
  const int64_t N = d->N;
  if (d->imp && (m->numeric || !h.full)) return fail(PLSPM_ERR_INVALID, "imputation needs the metric estimator and full tile sets");
  const size_t n_out = d->imp ? d->imp->base->h.n_out() : h.n_out();  // (rows of the BASE model when the handle is augmented)
  if (idx)
    for (int64_t e = 0; e < rep_count * N; ++e)
      if (idx[e] < 0 || idx[e] >= N) return fail(PLSPM_ERR_INVALID, "resample index out of range");
  const bool vote = !h.full && !m->numeric;
  const int64_t nb_max = bootstrap_batch_size(m, N, n_out, idx != nullptr, d->gram_mma, d->i8_colsum, d->Npad, d->n_zcols,
                                              d->Z8 != nullptr, d->sm_count, rep_count);
  BatchPlan bp;
  if (int rc = plan_batch(d, nb_max, bp)) return rc;
  const BatchBuffers bb = layout_batch(d, nb_max, bp, true, idx != nullptr, out_is_device != 0, false, false);
  // Batches are software-pipelined: the kernels of batch k + 1 are enqueued (into a second workspace) BEFORE the host
  // waits for batch k, so the device never idles while the host reads statuses and prepares the next launch -- the
  // 0.15 ms (one process) to 0.45 ms (eight processes sharing a host) per batch that kept the 8-GPU run at 0.96.
  // The host-driven non-metric path synchronises inside a batch anyway and stays serial.
  const int64_t n_batches = (rep_count + nb_max - 1) / nb_max;
  static const bool serial_env = getenv("PLSPM_PIPELINE") && std::string(getenv("PLSPM_PIPELINE")) == "0";
  const bool pipelined = n_batches > 1 && !m->numeric && !serial_env;
  const size_t ws_stride = align_up(bb.total, 4096);
  if (int rc = ws_reserve(d, pipelined ? 2 * ws_stride : bb.total)) return rc;
  cudaStream_t st = d->stream;
  PinnedStage stage((size_t)2 * (2 * nb_max + 2) * 4);
  if (!stage.ints) return fail(PLSPM_ERR_NOMEM, "pinned staging buffer");
  static const bool counts_split = getenv("PLSPM_COUNTS") && std::string(getenv("PLSPM_COUNTS")) == "split";
  static const bool sync_split = getenv("PLSPM_SYNC") && std::string(getenv("PLSPM_SYNC")) == "split";  // A/B: round 1's three syncs

  size_t timer_base = 0;  // stage-timer records of this call already collected (marks below are absolute)
  struct Batch {
    int64_t b0 = 0, nb = 0;
    size_t ws_off = 0;
    char* base = nullptr;
    uint32_t* cnt = nullptr;
    int32_t* idx_dev = nullptr;
    double* rows = nullptr;
    int *h_ovf = nullptr, *h_status = nullptr, *h_iters = nullptr;
    bool have_counts = false, used_i8 = false, used_fast_vote = false;  // routes the batch was ENQUEUED with (a later batch may be in flight when an earlier one switches them off)
    cudaEvent_t done = nullptr;
    size_t timer_mark = 0;  // stage-timer records up to here belong to batches <= this one
  } slot[2];
  for (int k = 0; k < 2; ++k) {
    slot[k].ws_off = (pipelined && k) ? ws_stride : 0;
    slot[k].base = (char*)d->ws.ptr + slot[k].ws_off;
    slot[k].h_ovf = stage.ints + (size_t)k * (2 * nb_max + 2);
    slot[k].h_status = slot[k].h_ovf + 2;
    slot[k].h_iters = slot[k].h_status + nb_max;
    if (cudaEventCreateWithFlags(&slot[k].done, cudaEventDisableTiming) != cudaSuccess) return fail(PLSPM_ERR_CUDA, "cudaEventCreate failed");
  }
  struct EventGuard { Batch* s; ~EventGuard() { for (int k = 0; k < 2; ++k) if (s[k].done) cudaEventDestroy(s[k].done); } } guard{slot};

  // the [nb x N] uint32 multiplicity table: what the fp64 kernels, the legacy routes and the exact redo read.  The
  // tensor-core routes take their int8 tile images straight from resample_images_kernel and never build it.
  auto ensure_counts = [&](Batch& B) -> int {
    if (B.have_counts) return 0;
    CK(cudaMemsetAsync(B.cnt, 0, (size_t)B.nb * N * 4, st));
    const int64_t threads = ((N + 3) / 4) * B.nb;
    d->timer.begin(ST_COUNTS, st);
    counts_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(B.cnt, B.idx_dev, N, B.nb, rep_begin + B.b0, seed);
    d->timer.end(st);
    CK(cudaGetLastError());
    B.have_counts = true;
    return 0;
  };
  // overflow flag, statuses, iteration counts and (host output) the rows of a batch come back together, behind one event
  auto read_back = [&](Batch& B) -> int {
    if (sync_split) {
      CK(cudaMemcpyAsync(B.h_ovf, B.base + bb.ovf, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      CK(cudaMemcpyAsync(B.h_status, B.base + bb.status, (size_t)B.nb * 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
    }
    CK(cudaMemcpyAsync(B.h_ovf, B.base + bb.ovf, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(B.h_status, B.base + bb.status, (size_t)B.nb * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(B.h_iters, B.base + bb.iters, (size_t)B.nb * 4, cudaMemcpyDeviceToHost, st));
    if (!out_is_device)
      CK(cudaMemcpyAsync(out + (size_t)B.b0 * n_out, B.rows, (size_t)B.nb * n_out * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaEventRecord(B.done, st));
    B.timer_mark = timer_base + d->timer.recs.size();
    return 0;
  };
  auto solve_batch = [&](Batch& B) -> int {
    d->ws_off = B.ws_off;
    if (m->numeric) return run_batch_num(d, B.nb, B.cnt, bb, scheme, tol, max_iter, bp, B.rows, false);
    return run_batch(d, B.nb, B.cnt, bb, scheme, tol, max_iter, bp, B.rows, false);
  };
  // everything of one batch up to its read-back, without waiting for anything
  auto enqueue = [&](Batch& B, int64_t b0) -> int {
    B.b0 = b0;
    B.nb = std::min(nb_max, rep_count - b0);  // (a short last batch reuses the plan and the workspace layout of a full one)
    B.have_counts = false;
    B.cnt = (uint32_t*)(B.base + bb.counts);
    B.idx_dev = nullptr;
    B.rows = out_is_device ? out + (size_t)b0 * n_out : (double*)(B.base + bb.out);
    const int64_t nb = B.nb;
    CK(cudaMemsetAsync(B.base + bb.ovf, 0, 8, st));
    if (idx) {
      B.idx_dev = (int32_t*)(B.base + bb.idx);
      CK(cudaMemcpyAsync(B.idx_dev, idx + b0 * N, (size_t)nb * N * 4, cudaMemcpyHostToDevice, st));
    }
    d->img_ready = false;
    d->c8img_ovr = d->c8vote_ovr = nullptr;
    const bool fuse_images = !counts_split && !m->numeric && d->gram_mma && d->i8_colsum && (!vote || (d->mma_vote && d->fast_vote));
    ImagePrefetch* pf = d->prefetch;
    if (fuse_images && pf && pf->valid && b0 == 0 && !idx && pf->nb == nb && pf->rep_begin == rep_begin && pf->seed == seed &&
        (pf->c8vote != nullptr) == vote) {
      // generated on the side stream while X was uploading: wait for it, adopt its overflow flag
      CK(cudaStreamWaitEvent(st, pf->done, 0));
      CK(cudaMemcpyAsync(B.base + bb.ovf, pf->ovf, 4, cudaMemcpyDeviceToDevice, st));
      d->c8img_ovr = pf->c8img;
      d->c8vote_ovr = pf->c8vote;
      d->img_ready = true;
    } else if (fuse_images) {
      const int64_t n_pad = (N + 127) / 128 * 128;
      const int64_t n_ranges = (n_pad + RI_MAX_ROWS - 1) / RI_MAX_ROWS;
      const int64_t range_rows = ((n_pad + n_ranges - 1) / n_ranges + 127) / 128 * 128;
      const int n_groups = (int)((nb + 511) / 512);
      CK(ensure_smem(resample_images_kernel, (size_t)(range_rows)));
      d->timer.begin(ST_COUNTS, st);
      resample_images_kernel<<<dim3((unsigned)n_groups * 512, (unsigned)n_ranges), RI_THREADS, (size_t)range_rows, st>>>(
          B.idx_dev, N, nb, rep_begin + b0, seed, range_rows, n_groups, vote ? (int)((nb + 127) / 128) : 0, d->n_chunks64,
          (uint8_t*)(B.base + bb.c8img), (uint8_t*)(B.base + bb.c8vote), (int*)(B.base + bb.ovf));
      d->timer.end(st);
      CK(cudaGetLastError());
      d->img_ready = true;
    } else if (int rc = ensure_counts(B)) {
      return rc;
    }
    B.used_i8 = d->i8_colsum;
    B.used_fast_vote = vote && d->fast_vote;
    if (int rc = solve_batch(B)) return rc;
    d->img_ready = false;
    d->c8img_ovr = d->c8vote_ovr = nullptr;
    return read_back(B);
  };
  // wait for a batch, run the (rare) redo paths, hand statuses / iteration counts to the caller
  auto finish = [&](Batch& B) -> int {
    CK(cudaEventSynchronize(B.done));
    const int64_t nb = B.nb;
    if (B.used_i8 && *B.h_ovf) {
      // a multiplicity above 127 does not fit the int8 operand of the tensor-core routes: redo the batch in fp64
      CK(cudaMemsetAsync(B.base + bb.ovf, 0, 8, st));
      d->i8_colsum = false;
      d->img_ready = false;
      d->c8img_ovr = d->c8vote_ovr = nullptr;
      if (int rc = ensure_counts(B)) return rc;
      B.used_i8 = false;
      B.used_fast_vote = vote && d->fast_vote;
      if (int rc = solve_batch(B)) return rc;
      if (int rc = read_back(B)) return rc;
      CK(cudaEventSynchronize(B.done));
    }
    if (B.used_fast_vote) {
      // replicates whose low-precision sign vote was undecided are redone with exact fp64 cross moments
      std::vector<int> redo;
      for (int64_t r = 0; r < nb; ++r)
        if (B.h_status[r] == STATUS_AMBIGUOUS) redo.push_back((int)r);
      if (!redo.empty()) {
        int* map_dev = (int*)(B.base + bb.rep_map);
        CK(cudaMemcpyAsync(map_dev, redo.data(), redo.size() * 4, cudaMemcpyHostToDevice, st));
        if (int rc = ensure_counts(B)) return rc;
        d->ws_off = B.ws_off;
        if (int rc = redo_exact(d, (int64_t)redo.size(), map_dev, B.cnt, bb, scheme, tol, max_iter, bp, B.rows)) return rc;
        g_redo_count += (int64_t)redo.size();
        if ((int64_t)redo.size() * 2 > nb) d->fast_vote = false;  // this data does not suit the fp16 vote
        if (int rc = read_back(B)) return rc;
        CK(cudaEventSynchronize(B.done));  // (redo.data() is pageable: the copy above has completed by now)
      }
    }
    if (iters) std::memcpy(iters + B.b0, B.h_iters, (size_t)nb * 4);
    if (status) std::memcpy(status + B.b0, B.h_status, (size_t)nb * 4);
    const size_t n_done = std::min(B.timer_mark - timer_base, d->timer.recs.size());
    d->timer.collect(n_done);
    timer_base += n_done;
    return 0;
  };

  int rc = enqueue(slot[0], 0);
  for (int64_t k = 0; k < n_batches && rc == 0; ++k) {
    Batch& cur = slot[pipelined ? (k & 1) : 0];
    if (pipelined && k + 1 < n_batches) rc = enqueue(slot[(k + 1) & 1], (k + 1) * nb_max);
    if (rc == 0) rc = finish(cur);
    if (rc == 0 && !pipelined && k + 1 < n_batches) rc = enqueue(slot[0], (k + 1) * nb_max);
  }
  d->ws_off = 0;
  if (rc) {
    cudaStreamSynchronize(st);
    d->timer.collect();
    return rc;
  }
  return PLSPM_OK;
}

int plspm_bootstrap_host(const plspm_model* m, const double* X, int64_t N, int64_t ld, int32_t scheme, double tol,
                         int32_t max_iter, int64_t rep_begin, int64_t rep_count, uint64_t seed, const int32_t* idx,
                         double* out, int32_t* status, int32_t* iters) {
  static const bool trace = getenv("PLSPM_TRACE") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  const int64_t miss0 = g_pool.misses.load(), ev0 = g_pool.evictions.load();
  plspm_data* d = nullptr;
  ImagePrefetch pf;
  if (m && X && !idx && out) prefetch_images(m, N, rep_begin, rep_count, seed, pf);
  int rc = plspm_data_create(m, X, N, ld, 0, &d);
  if (rc) {
    if (pf.done) cudaEventSynchronize(pf.done);
    prefetch_release(pf);
    return rc;
  }
  d->prefetch = &pf;
  const auto t1 = std::chrono::steady_clock::now();
  rc = plspm_bootstrap(m, d, scheme, tol, max_iter, rep_begin, rep_count, seed, idx, out, 0, status, iters);
  const auto t2 = std::chrono::steady_clock::now();
  d->prefetch = nullptr;
  if (pf.done) cudaEventSynchronize(pf.done);  // (an unused prefetch may still be running)
  prefetch_release(pf);
  plspm_data_destroy(d);
  if (trace) {
    const auto t3 = std::chrono::steady_clock::now();
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::milli>(b - a).count();
    };
    fprintf(stderr, "[plspm trace] bootstrap_host: create %.2f ms, bootstrap %.2f ms, destroy %.2f ms; pool misses %lld evictions %lld cached %.2f GB\n",
            ms(t0, t1), ms(t1, t2), ms(t2, t3), (long long)(g_pool.misses.load() - miss0), (long long)(g_pool.evictions.load() - ev0),
            (double)g_pool.cached / 1073741824.0);
  }
  return rc;
}

int plspm_data_set_imputation(plspm_data* d, const plspm_model* base, const int8_t* has_missing) {
  if (!d || !base || !has_missing) return fail(PLSPM_ERR_INVALID, "null argument");
  const HostModel& ha = d->model->h;
  const HostModel& hb = base->h;
  if (d->imp) return fail(PLSPM_ERR_INVALID, "imputation already set on this handle");
  if (!ha.full || !hb.full || d->model->numeric || base->numeric)
    return fail(PLSPM_ERR_INVALID, "imputation needs metric models with the full tile set (PLSPM_TILES_FULL)");
  if (ha.L != hb.L) return fail(PLSPM_ERR_INVALID, "augmented and base model have different latent variables");
  std::vector<int> ax(hb.Ppad, -1), am(hb.Ppad, -1);
  int src = 0;
  for (int l = 0; l < hb.L; ++l) {
    int n_miss = 0;
    for (int r = 0; r < hb.lv_k[l]; ++r, ++src) {
      // base source column `src` = column r of block l; has_missing is in base source order
      const int pb = hb.lv_off[l] + r;
      ax[pb] = ha.lv_off[l] + r;
      if (has_missing[src]) am[pb] = ha.lv_off[l] + hb.lv_k[l] + n_miss++;
    }
    if (ha.lv_k[l] != hb.lv_k[l] + n_miss)
      return fail(PLSPM_ERR_INVALID, "augmented block must be the base block followed by one indicator per column with missing values");
  }
  ImputeCtx* ic = new ImputeCtx();
  ic->base = base;
  d->imp = ic;
  CK(g_pool.alloc((void**)&ic->ax, (size_t)hb.Ppad * 4));
  CK(g_pool.alloc((void**)&ic->am, (size_t)hb.Ppad * 4));
  CK(g_pool.alloc((void**)&ic->mu_base, (size_t)hb.Ppad * 8));
  CK(cudaMemcpyAsync(ic->ax, ax.data(), (size_t)hb.Ppad * 4, cudaMemcpyHostToDevice, d->stream));
  CK(cudaMemcpyAsync(ic->am, am.data(), (size_t)hb.Ppad * 4, cudaMemcpyHostToDevice, d->stream));
  gather_kernel<<<(hb.Ppad + 127) / 128, 128, 0, d->stream>>>(d->mu, ic->ax, hb.Ppad, ic->mu_base);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(d->stream));
  return PLSPM_OK;
}

int plspm_resample_indices(uint64_t seed, int64_t replicate, int64_t N, int32_t* idx_out) {
  if (!idx_out || N < 1 || N >= ((int64_t)1 << 32)) return fail(PLSPM_ERR_INVALID, "bad arguments");
  int32_t* dev = nullptr;
  CK(cudaMalloc((void**)&dev, (size_t)N * 4));
  const int64_t groups = (N + 3) / 4;
  indices_kernel<<<(unsigned)((groups + 255) / 256), 256>>>(dev, N, (uint64_t)replicate, seed);
  cudaError_t e = cudaMemcpy(idx_out, dev, (size_t)N * 4, cudaMemcpyDeviceToHost);
  cudaFree(dev);
  if (e != cudaSuccess) return fail(PLSPM_ERR_CUDA, cudaGetErrorString(e));
  return PLSPM_OK;
}

int plspm_profile_reset(void) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (int i = 0; i < ST_N; ++i) { g_prof.ms[i] = 0; g_prof.launches[i] = 0; }
  return PLSPM_OK;
}
int plspm_profile_get(double* ms, int64_t* launches) {
  std::lock_guard<std::mutex> lk(g_prof.mu);
  for (int i = 0; i < ST_N; ++i) {
    if (ms) ms[i] = g_prof.ms[i];
    if (launches) launches[i] = g_prof.launches[i];
  }
  return PLSPM_OK;
}

int plspm_redo_count(int64_t* count) {
  if (!count) return fail(PLSPM_ERR_INVALID, "null argument");
  *count = g_redo_count.load();
  return PLSPM_OK;
}

int plspm_pool_trim(void) {
  g_pool.trim();
  return PLSPM_OK;
}
int plspm_pool_set_limit(int64_t bytes) {
  g_pool.set_limit(bytes < 0 ? 0 : (size_t)bytes);
  return PLSPM_OK;
}

int plspm_host_alloc(void** ptr, int64_t bytes) {
  CK(cudaMallocHost(ptr, (size_t)bytes));
  return PLSPM_OK;
}
int plspm_host_free(void* ptr) {
  CK(cudaFreeHost(ptr));
  return PLSPM_OK;
}

}  // extern "C"
