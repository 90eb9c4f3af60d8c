// Host-side lowering of a PLS-PM model (blocks, modes, path matrix) into the flat
// tables the sm_100a kernels consume.  Pure C++ (no CUDA), shared by the CUDA
// library and by the test-only host emulation of the per-replicate solver.
//
// Reference counterparts: Config.odm (config.py:140-144), Config.mode/mvs
// (config.py:146-166), Structure.path (config.py:51-58); InnerModel's effect rows
// (inner_model.py:50-60).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace plspm {

constexpr int SLOT = 8;             // doubles per slot (one 64-byte line)
constexpr int TILE = SLOT * SLOT;   // doubles per Gram tile

enum Scheme { SCHEME_CENTROID = 0, SCHEME_FACTORIAL = 1, SCHEME_PATH = 2 };
enum Mode { MODE_A = 0, MODE_B = 1 };
enum Status { STATUS_OK = 0, STATUS_NOT_CONVERGED = 1, STATUS_SINGULAR = 2,
              STATUS_AMBIGUOUS = 16 /* internal: low-precision sign vote undecided, redo exactly */ };
enum TilePolicy { TILES_AUTO = 0, TILES_FULL = 1, TILES_SPARSE = 2 };

// Raw-pointer view; the pointers are device pointers in the CUDA library and host
// pointers in the emulation.
struct ModelView {
  int L, P, Ppad, ns, scaled, full, n_tiles, n_tg, n_pairs, n_v, n_eff, max_deg, ws_doubles, kmax;
  int ng, n_cross, n_tg_cross;              // cross-moment tiles: ns x ng tiles of 8 MVs x 8 LVs (sparse tile sets)
  const int *lv_off, *lv_k, *lv_mode, *col_lv, *col_src;
  const int8_t* path;                       // [L*L] path[i*L+j]==1 : j -> i
  const int *tile_sa, *tile_sb, *tile_of;   // tile list and ns*ns lookup (see tile_of encoding)
  const int* lane_tile;                     // [n_tg*32] tile id handled by lane (or -1)
  const int *pair_l, *pair_j, *pair_voff, *lv_pair_begin;
  const int *eff_from, *eff_to;
  const int *chol_b_off;
  const int *pred_begin, *pred_idx, *succ_begin, *succ_idx;
  const int8_t* omega;                      // [L*L] 1: the Gram tiles of LV pair (i, j) are computed
};

#if defined(__CUDACC__)
#define PLSPM_HOST_DEVICE __host__ __device__
#else
#define PLSPM_HOST_DEVICE
#endif
// Per-LV scratch of the inner regressions: A [deg x deg] | beta [deg] | four vectors for the minimum-norm fallback
PLSPM_HOST_DEVICE inline int ols_scratch_doubles(int max_deg) { return max_deg * max_deg + 5 * max_deg; }
// Mode-B block of k manifest variables: Cholesky factor (or, rank deficient, the matrix itself) [k x k] | flag |
// solution + three vectors of the minimum-norm fallback
PLSPM_HOST_DEVICE inline int mode_b_scratch_doubles(int k) { return k * k + 1 + 4 * k; }

struct HostModel {
  int L = 0, P = 0, Ppad = 0, ns = 0, scaled = 0, full = 0;
  int n_tiles = 0, n_tg = 0, n_pairs = 0, n_v = 0, n_eff = 0, max_deg = 0, ws_doubles = 0, kmax = 0;
  // Sparse tile sets do not hold cov(x_p, score_l) for every (p, l), which the sign vote needs
  // (quirk Q6); a second streaming pass accumulates those P x L cross moments in tiles of
  // 8 MVs x 8 LVs: tile c = sa * ng + g covers slot sa and LVs 8g..8g+7.
  int ng = 0, n_cross = 0, n_tg_cross = 0;
  std::vector<int> lv_off, lv_k, lv_mode;   // [L+1] padded column offset, [L] block size, [L] mode
  std::vector<int> col_lv, col_src;         // [Ppad] LV of a padded column (-1 = padding), source column
  std::vector<int> src_col;                 // [P] padded column of source column p
  std::vector<int8_t> path;
  // Gram tiles: tile t covers slot pair (tile_sa[t] >= tile_sb[t]); element (r,c) of the tile is
  // sum_i c_i x[i, 8*sa+r] x[i, 8*sb+c].  tile_of[a*ns+b] = t if stored as (a,b), -(t+2) if stored
  // transposed, -1 if the slot pair is not computed.
  std::vector<int> tile_sa, tile_sb, tile_of;
  // Tile -> (tile group, lane) packing: tiles of one row slot stay in one warp so that the row
  // operand is a shared-memory broadcast; lane_tile[g*32+lane] = tile id or -1.
  std::vector<int> lane_tile;
  // directed LV pairs (l <- j) whose block covariance the iteration needs; sorted by l, the
  // diagonal pair (l <- l) first.  V[pair_voff[d] + r], r < K_l, receives (S_lj w_j)[r].
  std::vector<int> pair_l, pair_j, pair_voff, lv_pair_begin;
  std::vector<int> eff_from, eff_to;        // structurally reachable (from, to) pairs, reference row order
  std::vector<int> chol_b_off;              // [L] offset of the Mode-B Cholesky factor in the workspace, or -1
  std::vector<int> pred_begin, pred_idx, succ_begin, succ_idx;
  std::vector<int8_t> omega;                // [L*L] LV pairs covered by the tile set

  int n_out() const { return 2 * P + L + 2 * n_eff; }
  // shared-memory doubles the solver needs (see solver_core.h layout)
  size_t solver_smem_doubles() const { return (size_t)4 * Ppad + n_v + 4 * (size_t)L + 3 * (size_t)L * L + 40 + 2 * ((L + 1) / 2) + 8; }
  // solve_replicate (solver_core.h) keeps TWO L x L arrays (the path coefficients reuse the inner weights' storage, the
  // total effects the score correlations') and THREE Ppad vectors (no separate copy of the previous weights): 10 KB less
  // for c3, which is seven resident CTAs per SM instead of five
  size_t solver_core_smem_doubles() const { return solver_smem_doubles() - (size_t)L * L - (size_t)Ppad; }
  ModelView host_view() const;
};

// Returns 0 on success, else fills err.
int build_model(int L, const int32_t* block_sizes, const int8_t* modes, const int8_t* path, int scaled,
                int tile_policy, HostModel& m, std::string& err);

}  // namespace plspm
