// See plspm_model.h.
#include "plspm_model.h"

#include <algorithm>
#include <cstdlib>
#include <set>
#include <utility>

namespace plspm {

ModelView HostModel::host_view() const {
  ModelView v;
  v.L = L; v.P = P; v.Ppad = Ppad; v.ns = ns; v.scaled = scaled; v.full = full;
  v.n_tiles = n_tiles; v.n_tg = n_tg; v.n_pairs = n_pairs; v.n_v = n_v; v.n_eff = n_eff;
  v.max_deg = max_deg; v.ws_doubles = ws_doubles; v.kmax = kmax;
  v.ng = ng; v.n_cross = n_cross; v.n_tg_cross = n_tg_cross;
  v.lv_off = lv_off.data(); v.lv_k = lv_k.data(); v.lv_mode = lv_mode.data();
  v.col_lv = col_lv.data(); v.col_src = col_src.data(); v.path = path.data();
  v.tile_sa = tile_sa.data(); v.tile_sb = tile_sb.data(); v.tile_of = tile_of.data();
  v.lane_tile = lane_tile.data();
  v.pair_l = pair_l.data(); v.pair_j = pair_j.data(); v.pair_voff = pair_voff.data();
  v.lv_pair_begin = lv_pair_begin.data();
  v.eff_from = eff_from.data(); v.eff_to = eff_to.data(); v.chol_b_off = chol_b_off.data();
  v.pred_begin = pred_begin.data(); v.pred_idx = pred_idx.data();
  v.succ_begin = succ_begin.data(); v.succ_idx = succ_idx.data();
  v.omega = omega.data();
  return v;
}

int build_model(int L, const int32_t* block_sizes, const int8_t* modes, const int8_t* path, int scaled,
                int tile_policy, HostModel& m, std::string& err) {
  if (L < 1) { err = "model needs at least one latent variable"; return 1; }
  m = HostModel();
  m.L = L;
  m.scaled = scaled ? 1 : 0;
  m.lv_off.assign(L + 1, 0);
  m.lv_k.resize(L);
  m.lv_mode.resize(L);
  int slot = 0, P = 0;
  for (int l = 0; l < L; ++l) {
    if (block_sizes[l] < 1) { err = "every latent variable needs at least one manifest variable"; return 1; }
    if (modes[l] != MODE_A && modes[l] != MODE_B) { err = "mode must be 0 (A) or 1 (B)"; return 1; }
    m.lv_k[l] = block_sizes[l];
    m.kmax = std::max(m.kmax, (int)block_sizes[l]);
    m.lv_mode[l] = modes[l];
    m.lv_off[l] = slot * SLOT;
    slot += (block_sizes[l] + SLOT - 1) / SLOT;
    P += block_sizes[l];
  }
  m.lv_off[L] = slot * SLOT;
  m.ns = slot;
  m.Ppad = slot * SLOT;
  m.P = P;
  m.col_lv.assign(m.Ppad, -1);
  m.col_src.assign(m.Ppad, -1);
  m.src_col.assign(P, -1);
  int src = 0;
  for (int l = 0; l < L; ++l)
    for (int r = 0; r < m.lv_k[l]; ++r) {
      int c = m.lv_off[l] + r;
      m.col_lv[c] = l;
      m.col_src[c] = src;
      m.src_col[src] = c;
      ++src;
    }

  m.path.assign(path, path + (size_t)L * L);
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      int8_t v = m.path[(size_t)i * L + j];
      if (v != 0 && v != 1) { err = "path matrix entries must be 0 or 1"; return 1; }
      if (v == 1 && j >= i) { err = "path matrix must be strictly lower triangular"; return 1; }
    }
  // predecessor / successor lists
  m.pred_begin.assign(L + 1, 0);
  m.succ_begin.assign(L + 1, 0);
  for (int i = 0; i < L; ++i) {
    m.pred_begin[i] = (int)m.pred_idx.size();
    for (int j = 0; j < L; ++j)
      if (m.path[(size_t)i * L + j]) m.pred_idx.push_back(j);
    m.max_deg = std::max(m.max_deg, (int)m.pred_idx.size() - m.pred_begin[i]);
    m.succ_begin[i] = (int)m.succ_idx.size();
    for (int k = 0; k < L; ++k)
      if (m.path[(size_t)k * L + i]) m.succ_idx.push_back(k);
  }
  m.pred_begin[L] = (int)m.pred_idx.size();
  m.succ_begin[L] = (int)m.succ_idx.size();

  // LV pairs the iteration touches: diagonal, edges, co-parents (inner OLS / path scheme).
  std::set<std::pair<int, int>> und;  // (hi, lo), hi > lo
  for (int i = 0; i < L; ++i) {
    for (int a = m.pred_begin[i]; a < m.pred_begin[i + 1]; ++a) {
      int j = m.pred_idx[a];
      und.insert({std::max(i, j), std::min(i, j)});
      for (int b = a + 1; b < m.pred_begin[i + 1]; ++b) {
        int k = m.pred_idx[b];
        und.insert({std::max(j, k), std::min(j, k)});
      }
    }
  }
  std::vector<std::vector<int>> nbr(L);
  for (auto& pr : und) {
    nbr[pr.first].push_back(pr.second);
    nbr[pr.second].push_back(pr.first);
  }
  m.lv_pair_begin.assign(L + 1, 0);
  int voff = 0;
  for (int l = 0; l < L; ++l) {
    m.lv_pair_begin[l] = (int)m.pair_l.size();
    std::sort(nbr[l].begin(), nbr[l].end());
    auto add = [&](int j) {
      m.pair_l.push_back(l);
      m.pair_j.push_back(j);
      m.pair_voff.push_back(voff);
      voff += m.lv_k[l];
    };
    add(l);
    for (int j : nbr[l]) add(j);
  }
  m.lv_pair_begin[L] = (int)m.pair_l.size();
  m.n_pairs = (int)m.pair_l.size();
  m.n_v = voff;

  // Gram tiles.  Sparse = only the slot pairs of LV pairs the iteration touches, plus a P x L
  // cross-moment pass for the sign vote; AUTO picks whichever costs fewer fp64 FMAs per row.
  m.ng = (L + SLOT - 1) / SLOT;
  m.n_cross = m.ns * m.ng;
  m.n_tg_cross = (m.n_cross + 31) / 32;
  std::set<std::pair<int, int>> sparse_tiles, tiles;  // (sa, sb) with sa >= sb
  {
    auto add_lv_pair = [&](int l, int j) {
      for (int a = m.lv_off[l] / SLOT; a < m.lv_off[l + 1] / SLOT; ++a)
        for (int b = m.lv_off[j] / SLOT; b < m.lv_off[j + 1] / SLOT; ++b)
          sparse_tiles.insert({std::max(a, b), std::min(a, b)});
    };
    for (int l = 0; l < L; ++l) add_lv_pair(l, l);
    for (auto& pr : und) add_lv_pair(pr.first, pr.second);
  }
  const double cost_full = 0.5 * m.ns * (m.ns + 1);
  const double cost_sparse = (double)sparse_tiles.size() + 1.15 * m.n_cross;
  bool full = (tile_policy == TILES_FULL) || (tile_policy == TILES_AUTO && cost_sparse > 0.8 * cost_full);
  m.full = full ? 1 : 0;
  m.tile_of.assign((size_t)m.ns * m.ns, -1);
  if (full) {
    for (int a = 0; a < m.ns; ++a)
      for (int b = 0; b <= a; ++b) tiles.insert({a, b});
  } else {
    tiles = sparse_tiles;
  }
  for (auto& t : tiles) {
    int id = (int)m.tile_sa.size();
    m.tile_sa.push_back(t.first);
    m.tile_sb.push_back(t.second);
    m.tile_of[(size_t)t.first * m.ns + t.second] = id;
    if (t.first != t.second) m.tile_of[(size_t)t.second * m.ns + t.first] = -(id + 2);
  }
  m.omega.assign((size_t)L * L, full ? 1 : 0);
  if (!full) {
    for (int l = 0; l < L; ++l) m.omega[(size_t)l * L + l] = 1;
    for (auto& pr : und) m.omega[(size_t)pr.first * L + pr.second] = m.omega[(size_t)pr.second * L + pr.first] = 1;
  }
  m.n_tiles = (int)m.tile_sa.size();
  // pack tiles into warps (tile groups of 32 lanes): chunks of one row slot, first-fit decreasing.
  // Diagonal tiles (sa == sb) go into groups of their own when that costs no extra group: such a warp
  // needs one operand per row and 36 of the 64 products (gram_kernel, "diag"); they are listed last so
  // that the shorter CTAs fill the tail of the launch.
  {
    const bool pack = !(getenv("PLSPM_TILE_PACK") && atoi(getenv("PLSPM_TILE_PACK")) == 0);
    auto pack_bins = [&](const std::vector<int>& ids) {
      std::vector<std::vector<int>> chunks;
      for (size_t t = 0; t < ids.size();) {
        size_t e = t;
        while (e < ids.size() && m.tile_sa[ids[e]] == m.tile_sa[ids[t]] && e - t < 32) ++e;
        chunks.emplace_back(ids.begin() + t, ids.begin() + e);
        t = e;
      }
      if (!pack) {  // natural order, 32 consecutive tiles per group (experiment switch)
        chunks.clear();
        for (size_t t = 0; t < ids.size(); t += 32)
          chunks.emplace_back(ids.begin() + t, ids.begin() + std::min(t + 32, ids.size()));
      }
      std::stable_sort(chunks.begin(), chunks.end(),
                       [](const std::vector<int>& a, const std::vector<int>& b) { return a.size() > b.size(); });
      std::vector<std::vector<int>> bins;
      for (auto& ch : chunks) {
        bool placed = false;
        for (auto& b : bins)
          if (b.size() + ch.size() <= 32) { b.insert(b.end(), ch.begin(), ch.end()); placed = true; break; }
        if (!placed) bins.push_back(ch);
      }
      return bins;
    };
    std::vector<int> all_ids, diag_ids, off_ids;
    for (int t = 0; t < m.n_tiles; ++t) {
      all_ids.push_back(t);
      (m.tile_sa[t] == m.tile_sb[t] ? diag_ids : off_ids).push_back(t);
    }
    std::vector<std::vector<int>> bins = pack_bins(all_ids);
    const bool split_ok = !(getenv("PLSPM_TILE_DIAG") && atoi(getenv("PLSPM_TILE_DIAG")) == 0);
    if (split_ok && pack && !off_ids.empty()) {
      std::vector<std::vector<int>> off_bins = pack_bins(off_ids), diag_bins;
      for (size_t t = 0; t < diag_ids.size(); t += 32)
        diag_bins.emplace_back(diag_ids.begin() + t, diag_ids.begin() + std::min(t + 32, diag_ids.size()));
      if (off_bins.size() + diag_bins.size() <= bins.size()) {
        bins = off_bins;
        bins.insert(bins.end(), diag_bins.begin(), diag_bins.end());
      }
    }
    m.n_tg = (int)bins.size();
    m.lane_tile.assign((size_t)m.n_tg * 32, -1);
    for (int g = 0; g < m.n_tg; ++g)
      for (size_t k = 0; k < bins[g].size(); ++k) m.lane_tile[(size_t)g * 32 + k] = bins[g][k];
  }

  // effect rows: (from, to) with a directed path from -> to, reference row order
  std::vector<char> reach((size_t)L * L, 0);  // reach[to*L+from]
  for (int i = 0; i < L; ++i)  // LVs are topologically ordered (strictly lower-triangular path)
    for (int a = m.pred_begin[i]; a < m.pred_begin[i + 1]; ++a) {
      int j = m.pred_idx[a];
      reach[(size_t)i * L + j] = 1;
      for (int f = 0; f < L; ++f)
        if (reach[(size_t)j * L + f]) reach[(size_t)i * L + f] = 1;
    }
  for (int f = 0; f < L; ++f)
    for (int t = 0; t < L; ++t)
      if (f != t && reach[(size_t)t * L + f]) {
        m.eff_from.push_back(f);
        m.eff_to.push_back(t);
      }
  m.n_eff = (int)m.eff_from.size();

  // per-replicate global workspace: Mode-B Cholesky factors + per-LV OLS scratch
  m.chol_b_off.assign(L, -1);
  int ws = 0;
  for (int l = 0; l < L; ++l)
    if (m.lv_mode[l] == MODE_B) {
      m.chol_b_off[l] = ws;
      ws += mode_b_scratch_doubles(m.lv_k[l]);
    }
  ws += L * ols_scratch_doubles(m.max_deg);
  m.ws_doubles = std::max(ws, 1);
  return 0;
}

}  // namespace plspm
