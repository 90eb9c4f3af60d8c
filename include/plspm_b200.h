/* plspm_b200 -- C ABI of the B200-native PLS-PM weight-estimation engine.
 *
 * The reference (GoogleCloudPlatform/plspm-python @ 37f4aaf) has no FFI: the narrowest
 * seams around its hot path are Python methods.  Each entry point below names the
 * reference interface it replaces (file:line under /root/reference).  The Python host
 * package (plspm-python_b200/plspm) binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions: plain pointers and sizes; the caller owns every output buffer; the
 * library owns only the opaque handles; functions return 0 on success and a non-zero
 * PLSPM_ERR_* code otherwise (plspm_last_error() gives the message); no exception
 * crosses the boundary.  Per-replicate failures are NOT errors of the call: they are
 * reported in status[] so the host can drop the replicate exactly like the reference's
 * bare `except: pass` (bootstrap.py:67-68).
 *
 * Column order of every [P]-sized array and of X: manifest variables grouped by latent
 * variable in PATH-MATRIX order (the ODM row order, config.py:140-144, weights.py:31,69).
 */
#ifndef PLSPM_B200_H
#define PLSPM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct plspm_model plspm_model; /* lowered Config: blocks, modes, path matrix */
typedef struct plspm_data plspm_data;   /* observation x manifest matrix resident in HBM */

enum { PLSPM_OK = 0, PLSPM_ERR_INVALID = 1, PLSPM_ERR_CUDA = 2, PLSPM_ERR_NOMEM = 3, PLSPM_ERR_UNSUPPORTED = 4 };
/* Scheme (scheme.py:57-63) and Mode (mode.py:64-69) */
enum { PLSPM_SCHEME_CENTROID = 0, PLSPM_SCHEME_FACTORIAL = 1, PLSPM_SCHEME_PATH = 2 };
enum { PLSPM_MODE_A = 0, PLSPM_MODE_B = 1 };
/* per-fit status: converged / "Could not converge after N iterations" (weights.py:185-186)
 * / non-finite moments in a Mode-B block or an inner regression (rank-deficient ones get the minimum-norm solution
 * the reference's lstsq / pinv return, mode.py:50-52, inner_model.py:76-77) */
enum { PLSPM_FIT_OK = 0, PLSPM_FIT_NOT_CONVERGED = 1, PLSPM_FIT_SINGULAR = 2 };
/* Gram tile policy: which manifest-variable cross moments the Gram kernel accumulates */
enum { PLSPM_TILES_AUTO = 0, PLSPM_TILES_FULL = 1, PLSPM_TILES_SPARSE = 2 };

int plspm_version(void);
const char* plspm_last_error(void);
int plspm_device_count(int32_t* count);
int plspm_set_device(int32_t device);

/* Replaces Config(path, scaled) + add_lv(...) + Config.odm / .mode / .path as consumed by
 * the hot path (config.py:84-123, 136-166, 178-201).  path is row-major L x L, path[i*L+j]
 * == 1 means LV j -> LV i (strictly lower triangular, config.py:112-115).  scaled is
 * Config's `scaled` flag (pooled scaling, config.py:301-303). */
int plspm_model_create(int32_t L, const int32_t* block_sizes, const int8_t* modes, const int8_t* path,
                       int32_t scaled, int32_t tile_policy, plspm_model** out);
void plspm_model_destroy(plspm_model* m);
/* info[16]: 0 L, 1 P, 2 padded P, 3 tiles, 4 tile groups, 5 LV pairs, 6 effect rows, 7 doubles per
 * bootstrap row (2P + L + 2*effects), 8 full tile set, 9 scaled, 10 cross-moment tiles, 11 numeric,
 * 12 pair-product columns of the tile set (the int8 Gram GEMM has 6x as many rows) */
int plspm_model_query(const plspm_model* m, int32_t* info);
/* on != 0 selects the reference's non-metric estimator for data whose manifest variables all carry
 * Scale.NUM or Scale.RAW (config.py:306-319 treatment, weights.py:73-133 `_NonmetricWeights`, mode.py
 * 31-42 / 54-61): columns standardised per (re)sample, no sign vote, stopping rule on the scores.  The
 * `scaled` flag of the model is ignored on this path (the reference ignores it too). */
int plspm_model_set_numeric(plspm_model* m, int32_t on);
/* (from, to) LV ids of the effect rows, in the row order of InnerModel.effects()
 * (inner_model.py:50-60).  Arrays of info[6] entries. */
int plspm_model_effects(const plspm_model* m, int32_t* from, int32_t* to);

/* Replaces Config.filter's column selection + the data side of Config.treat
 * (config.py:269, 299-305): uploads X (row-major N x P doubles, leading dimension ld, host or
 * device pointer) once and keeps it resident for any number of fits / bootstrap calls. */
/* A data handle depends only on the column layout (block sizes in path order) of the model it was created
 * with; plspm_fit / plspm_bootstrap accept it together with any model of the same layout.  The handle owns
 * the stream and the workspace of the call in flight: one call at a time per handle (hence non-const). */
int plspm_data_create(const plspm_model* m, const double* X, int64_t N, int64_t ld, int32_t x_is_device,
                      plspm_data** out);
void plspm_data_destroy(plspm_data* d);

/* Replaces Estimator.estimate -> WeightsCalculatorFactory.calculate (estimator.py:29-55,
 * weights.py:172-187: _MetricWeights.__init__/iterate/calculate, scheme.py, mode.py) plus the
 * InnerModel / OuterModel quantities derived from the scores (inner_model.py:66-83,
 * outer_model.py:24-34).  Host output pointers, any may be NULL:
 * weights[P], loadings[P], r_squared[L], paths[L*L] (row = to, col = from), total_effects[L*L],
 * crossloadings[P*L], scores[N*L].  iters = iterate() calls made, status = PLSPM_FIT_*. */
int plspm_fit(const plspm_model* m, plspm_data* d, int32_t scheme, double tol, int32_t max_iter,
              double* weights, double* loadings, double* r_squared, double* paths, double* total_effects,
              double* crossloadings, double* scores, int32_t* iters, int32_t* status);

/* Replaces BootstrapProcess.run (bootstrap.py:45-75) for global replicate ids
 * [rep_begin, rep_begin + rep_count).  idx == NULL: resample indices come from Philox4x32-10
 * keyed by (seed, global replicate id) (plspm_resample_indices gives the same stream);
 * otherwise idx is a HOST int32 [rep_count * N] matrix of injected indices (parity runs).
 * out: [rep_count * info[7]] doubles, each row = weights P | r_squared L | total effects E |
 * direct effects E | loadings P; a host pointer, or a device pointer if out_is_device.
 * status / iters: host int32 [rep_count].
 * Replicates run in batches sized by a 3 GB workspace budget (PLSPM_MAX_BATCH caps them); with several batches the
 * kernels of batch k + 1 are enqueued before the host waits for batch k (two workspaces).  The call returns after the
 * stream has drained. */
int plspm_bootstrap(const plspm_model* m, plspm_data* d, int32_t scheme, double tol, int32_t max_iter,
                    int64_t rep_begin, int64_t rep_count, uint64_t seed, const int32_t* idx, double* out,
                    int32_t out_is_device, int32_t* status, int32_t* iters);

/* Same call with a HOST observation matrix: upload + bootstrap + release in one entry point
 * (what a caller without a resident data handle pays end to end).  With idx == NULL the int8 multiplicity images of
 * the first batch are generated on a side stream while X is uploading. */
int plspm_bootstrap_host(const plspm_model* m, const double* X, int64_t N, int64_t ld, int32_t scheme, double tol,
                         int32_t max_iter, int64_t rep_begin, int64_t rep_count, uint64_t seed, const int32_t* idx,
                         double* out, int32_t* status, int32_t* iters);

/* Bootstrap on data with missing values (replaces the per-replicate re-imputation of bootstrap.py:57 ->
 * estimator.py:33 -> config.py:299-305 -> util.py:61-68, column means of the replicate's observed rows).
 * `d` must hold the AUGMENTED matrix [x0 | m] under an augmented model: block l = the base block's columns with
 * missing entries set to 0, followed by one 0/1 missing indicator per column of the block that has missing entries.
 * base: the model without indicators.  Both models need PLSPM_TILES_FULL and the metric estimator.
 * has_missing: host int8 [base P] in base column order.  After this call plspm_bootstrap(augmented model, d, ...)
 * returns rows of the BASE model (base info[7] doubles per replicate).  base must outlive d. */
int plspm_data_set_imputation(plspm_data* d, const plspm_model* base, const int8_t* has_missing);

/* The resample index stream of one replicate (replaces numpy.random.randint at
 * bootstrap.py:56, which is unseeded in the reference).  idx_out: host int32 [N]. */
int plspm_resample_indices(uint64_t seed, int64_t replicate, int64_t N, int32_t* idx_out);

/* Instrumentation of the library's own stream since the last reset: device milliseconds per
 * stage measured with CUDA events around every launch, and kernel launch counts.
 * ms[12] / launches[12]: 0 counts (resample multiplicities), 1 gram (fp64 Gram kernel), 2 chunk reduce, 3 solve,
 * 4 scores, 5 upload kernels, 6 column sums / int8 multiplicity images, 7 sign vote (fused tcgen05 kernel, or the
 * exact fp64 cross-moment pass), 8 score generation (legacy vote), 9 non-metric criterion pass, 10 integer Gram
 * (tcgen05 kernel; legacy: library GEMM + combine), 11 fp64 recombination of the integer Gram. */
int plspm_profile_reset(void);
int plspm_profile_get(double* ms, int64_t* launches);

/* Number of bootstrap replicates (since library load) whose low-precision tensor-core sign vote was
 * undecided and which were therefore redone with exact fp64 cross moments. */
int plspm_redo_count(int64_t* count);

/* Released device buffers (data handles, workspaces) are cached for the next call, up to a limit
 * (PLSPM_POOL_GB, default 8 GB).  plspm_pool_trim returns every cached buffer to the driver;
 * plspm_pool_set_limit changes the limit (0 = cache nothing). */
int plspm_pool_trim(void);
int plspm_pool_set_limit(int64_t bytes);

/* Pinned host memory for callers that want full-speed host<->device copies. */
int plspm_host_alloc(void** ptr, int64_t bytes);
int plspm_host_free(void* ptr);

#ifdef __cplusplus
}
#endif
#endif
