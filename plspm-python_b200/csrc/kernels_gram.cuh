// fp64 weighted Gram kernel (TMA ring, 8x8 register tiles), column sums, chunk reduction.
// Part of the single translation unit plspm_b200.cu (included there, in this order); see DESIGN.md §4.
#pragma once

// ------------------------------------------------------------------------------------------------
// weighted Gram kernel
// ------------------------------------------------------------------------------------------------
constexpr int GRAM_WARPS = 8;    // consumer warps per CTA, each = one (replicate, tile group) item
constexpr int GRAM_THREADS = GRAM_WARPS * 32;  // 2 warps per SM sub-partition: up to 255 registers per thread
constexpr int GRAM_MAX_STAGES = 8;

struct GramParams {
  const double* X;          // [N][Ppad]
  const uint32_t* counts;   // [nrep][N] or null (every row once)
  int64_t N;
  int Ppad, n_tiles, n_tg;
  const int *tile_sa, *tile_sb, *lane_tile;
  int64_t n_items;          // nrep * n_tg
  int n_chunks;
  int64_t chunk_rows;       // multiple of RT
  int RT, stages;
  double* G;                // [nrep][n_chunks][n_tiles*64]
  // cross-moment mode (template CROSS): tiles are (slot sa, LV group g) in natural order, the
  // column operand is the row's LV scores x~_i . wf_l times the multiplicity (per-warp scratch)
  int L, ng;
  const int *lv_off, *lv_k;
  const double* wf;         // [nrep][Ppad] final weights of every replicate
  const int* rep_map;       // optional: item / n_tg -> replicate (exact redo of selected replicates)
};

template <bool CROSS>
__global__ void __launch_bounds__(GRAM_THREADS, 1) gram_kernel(const GramParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[GRAM_MAX_STAGES];
  double* tiles = reinterpret_cast<double*>(smem_raw);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_groups = (p.n_items + GRAM_WARPS - 1) / GRAM_WARPS;
  const int chunk = (int)(blockIdx.x / n_groups);
  const int64_t group = blockIdx.x - (int64_t)chunk * n_groups;
  const int64_t item0 = group * GRAM_WARPS;
  const int n_active = (int)min((int64_t)GRAM_WARPS, p.n_items - item0);
  const int64_t r0 = (int64_t)chunk * p.chunk_rows, r1 = min(r0 + p.chunk_rows, p.N);
  const int n_rt = (int)((r1 - r0 + p.RT - 1) / p.RT);
  const size_t stage_doubles = (size_t)p.RT * p.Ppad;

  // ---- feeding the ring -------------------------------------------------------------------------
  // No producer warp (a ninth warp would put three warps on one SM sub-partition and cap every thread at
  // 168 registers; the 8x8 fp64 accumulator tile alone needs 128).  Thread 0 issues the first `stages`
  // row tiles; after that the LAST consumer warp to finish with a stage (shared-memory counter) refills it
  // through the TMA engine at once, so a tile is always requested stages-1 tile times ahead of its use
  // no matter how the warps drift apart.
  __shared__ int stage_done[GRAM_MAX_STAGES];
  auto issue_tile = [&](int tn) {
    const int st = tn % p.stages;
    const int64_t row = r0 + (int64_t)tn * p.RT;
    const uint32_t rows = (uint32_t)min((int64_t)p.RT, r1 - row);
    const uint32_t bytes = rows * (uint32_t)p.Ppad * 8u;
    mbar_arrive_expect_tx(&full_bar[st], bytes);
    bulk_g2s(tiles + (size_t)st * stage_doubles, p.X + row * p.Ppad, bytes, &full_bar[st]);
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      stage_done[s] = 0;
    }
    mbar_fence_init();
    for (int tn = 0; tn < n_rt && tn < p.stages; ++tn) issue_tile(tn);
  }
  __syncthreads();
  if (warp >= n_active) return;

  // ---- consumer warp: one (replicate, tile group); lane = one 8x8 tile ------------------------
  // items are tile-group major: the warps of a CTA work on the same tile group for 8 replicates, so a CTA
  // is (except at a group boundary) all "diagonal" or all "generic" warps -- see below
  const int64_t item = item0 + warp;
  const int64_t nrep_pos = p.n_items / p.n_tg;
  const int tg = (int)(item / nrep_pos);
  const int64_t rep_pos = item - (int64_t)tg * nrep_pos;
  const int64_t rep = p.rep_map ? (int64_t)p.rep_map[rep_pos] : rep_pos;
  int tile, sa, sb;
  if constexpr (CROSS) {
    // LV-group-major order: a warp covers (almost always) ONE group of 8 LVs and 32 row slots, so it
    // needs only that group's scores
    tile = tg * 32 + lane;
    if (tile >= p.n_tiles) tile = -1;
    const int ns = p.Ppad / SLOT;
    sb = tile >= 0 ? tile / ns : 0;          // LV group: "slot" sb of the score scratch row
    sa = tile >= 0 ? tile - sb * ns : 0;
  } else {
    tile = p.lane_tile[tg * 32 + lane];
    sa = tile >= 0 ? p.tile_sa[tile] : 0;
    sb = tile >= 0 ? p.tile_sb[tile] : 0;
  }
  const bool tile_ok = tile >= 0;
  // A tile group that holds only diagonal tiles (sa == sb; the model builder packs them together) needs one
  // operand per row and, by symmetry, 36 of the 64 products.
  const bool diag = !CROSS && __all_sync(0xffffffffu, !tile_ok || sa == sb);
  // 16-byte chunks of a slot are read in a lane-dependent rotated order so that the 32 LDS.128 of
  // a warp spread over all bank quads (slot stride 64 B would otherwise be a 16-way conflict).
  // This holds for the row operand too: a sparse tile group holds ~3 tiles per row slot, i.e. ~11
  // distinct row slots per warp (ncu: 8.0 wavefronts per unrotated xa LDS.128 vs 4.33 rotated; the
  // kernel is bound by shared-memory wavefronts, 94 % L1/TEX throughput, before the fp64 pipe).
  const int rot_a = (sa >> 1) & 3, rot_b = CROSS ? 0 : ((sb >> 1) & 3);
  int off_a[4], off_b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    off_a[k] = sa * SLOT + 2 * ((k + rot_a) & 3);
    off_b[k] = sb * SLOT + 2 * ((k + rot_b) & 3);
  }
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;

  const uint32_t* cnt_row = p.counts ? p.counts + rep * p.N : nullptr;
  auto load_counts = [&](int t_load) -> uint32_t {
    const int64_t row = r0 + (int64_t)t_load * p.RT;
    const int rows = (int)min((int64_t)p.RT, r1 - row);
    if (lane >= rows) return 0u;
    return cnt_row ? __ldg(cnt_row + row + lane) : 1u;
  };
  // shared-memory byte addresses of the lane's 2 x 4 16-byte operand chunks within a row
  uint32_t boff_a[4], boff_b[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    boff_a[k] = (uint32_t)off_a[k] * 8u;
    boff_b[k] = (uint32_t)off_b[k] * 8u;
  }
  auto lds128 = [](uint32_t addr, double& x, double& y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(addr));
  };
  auto load_row = [&](uint32_t row_addr, uint32_t b_addr, double (&xa)[8], double (&xb)[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      lds128(row_addr + boff_a[k], xa[2 * k], xa[2 * k + 1]);
      lds128(b_addr + boff_b[k], xb[2 * k], xb[2 * k + 1]);
    }
  };
  // The scaling and the 64 FMAs of a row are emitted as volatile asm so that the compiler keeps
  // them AFTER the (volatile) shared-memory loads of the NEXT row in program order: without this
  // the loads get sunk below the FMA block to save registers and the software pipeline is lost
  // (measured: 6.5 % of all issue slots stalled on the first DMUL of every row).
  // `scale` is a compile-time tag: rows of multiplicity 1 (58 % of the non-zero rows of a resample) are
  // listed first and skip the 8 multiplications
  auto accumulate = [&](auto scale, double (&xa)[8], double (&xb)[8], double c) {
    if constexpr (!CROSS && decltype(scale)::value)  // (the score scratch is already multiplied by the multiplicity)
    asm volatile(
        "mul.f64 %0, %0, %8;\n\tmul.f64 %1, %1, %8;\n\tmul.f64 %2, %2, %8;\n\tmul.f64 %3, %3, %8;\n\t"
        "mul.f64 %4, %4, %8;\n\tmul.f64 %5, %5, %8;\n\tmul.f64 %6, %6, %8;\n\tmul.f64 %7, %7, %8;"
        : "+d"(xb[0]), "+d"(xb[1]), "+d"(xb[2]), "+d"(xb[3]), "+d"(xb[4]), "+d"(xb[5]), "+d"(xb[6]), "+d"(xb[7])
        : "d"(c));
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile(
          "fma.rn.f64 %0, %8, %9, %0;\n\tfma.rn.f64 %1, %8, %10, %1;\n\tfma.rn.f64 %2, %8, %11, %2;\n\t"
          "fma.rn.f64 %3, %8, %12, %3;\n\tfma.rn.f64 %4, %8, %13, %4;\n\tfma.rn.f64 %5, %8, %14, %5;\n\t"
          "fma.rn.f64 %6, %8, %15, %6;\n\tfma.rn.f64 %7, %8, %16, %7;"
          : "+d"(acc[i][0]), "+d"(acc[i][1]), "+d"(acc[i][2]), "+d"(acc[i][3]), "+d"(acc[i][4]), "+d"(acc[i][5]),
            "+d"(acc[i][6]), "+d"(acc[i][7])
          : "d"(xa[i]), "d"(xb[0]), "d"(xb[1]), "d"(xb[2]), "d"(xb[3]), "d"(xb[4]), "d"(xb[5]), "d"(xb[6]),
            "d"(xb[7]));
  };
  auto accumulate_diag = [&](auto scale, double (&xa)[8], double (&xs)[8], double c) {  // xs = c * xa, upper triangle only
    if constexpr (!decltype(scale)::value) {
#pragma unroll
      for (int k = 0; k < 8; ++k) xs[k] = xa[k];
    } else
    asm volatile(
        "mul.f64 %0, %8, %16;\n\tmul.f64 %1, %9, %16;\n\tmul.f64 %2, %10, %16;\n\tmul.f64 %3, %11, %16;\n\t"
        "mul.f64 %4, %12, %16;\n\tmul.f64 %5, %13, %16;\n\tmul.f64 %6, %14, %16;\n\tmul.f64 %7, %15, %16;"
        : "=d"(xs[0]), "=d"(xs[1]), "=d"(xs[2]), "=d"(xs[3]), "=d"(xs[4]), "=d"(xs[5]), "=d"(xs[6]), "=d"(xs[7])
        : "d"(xa[0]), "d"(xa[1]), "d"(xa[2]), "d"(xa[3]), "d"(xa[4]), "d"(xa[5]), "d"(xa[6]), "d"(xa[7]), "d"(c));
    asm volatile("fma.rn.f64 %0, %8, %9, %0;\n\tfma.rn.f64 %1, %8, %10, %1;\n\tfma.rn.f64 %2, %8, %11, %2;\n\tfma.rn.f64 %3, %8, %12, %3;\n\tfma.rn.f64 %4, %8, %13, %4;\n\tfma.rn.f64 %5, %8, %14, %5;\n\tfma.rn.f64 %6, %8, %15, %6;\n\tfma.rn.f64 %7, %8, %16, %7;"
                 : "+d"(acc[0][0]), "+d"(acc[0][1]), "+d"(acc[0][2]), "+d"(acc[0][3]), "+d"(acc[0][4]), "+d"(acc[0][5]), "+d"(acc[0][6]), "+d"(acc[0][7])
                 : "d"(xa[0]), "d"(xs[0]), "d"(xs[1]), "d"(xs[2]), "d"(xs[3]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %7, %8, %0;\n\tfma.rn.f64 %1, %7, %9, %1;\n\tfma.rn.f64 %2, %7, %10, %2;\n\tfma.rn.f64 %3, %7, %11, %3;\n\tfma.rn.f64 %4, %7, %12, %4;\n\tfma.rn.f64 %5, %7, %13, %5;\n\tfma.rn.f64 %6, %7, %14, %6;"
                 : "+d"(acc[1][1]), "+d"(acc[1][2]), "+d"(acc[1][3]), "+d"(acc[1][4]), "+d"(acc[1][5]), "+d"(acc[1][6]), "+d"(acc[1][7])
                 : "d"(xa[1]), "d"(xs[1]), "d"(xs[2]), "d"(xs[3]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %6, %7, %0;\n\tfma.rn.f64 %1, %6, %8, %1;\n\tfma.rn.f64 %2, %6, %9, %2;\n\tfma.rn.f64 %3, %6, %10, %3;\n\tfma.rn.f64 %4, %6, %11, %4;\n\tfma.rn.f64 %5, %6, %12, %5;"
                 : "+d"(acc[2][2]), "+d"(acc[2][3]), "+d"(acc[2][4]), "+d"(acc[2][5]), "+d"(acc[2][6]), "+d"(acc[2][7])
                 : "d"(xa[2]), "d"(xs[2]), "d"(xs[3]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %5, %6, %0;\n\tfma.rn.f64 %1, %5, %7, %1;\n\tfma.rn.f64 %2, %5, %8, %2;\n\tfma.rn.f64 %3, %5, %9, %3;\n\tfma.rn.f64 %4, %5, %10, %4;"
                 : "+d"(acc[3][3]), "+d"(acc[3][4]), "+d"(acc[3][5]), "+d"(acc[3][6]), "+d"(acc[3][7])
                 : "d"(xa[3]), "d"(xs[3]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %4, %5, %0;\n\tfma.rn.f64 %1, %4, %6, %1;\n\tfma.rn.f64 %2, %4, %7, %2;\n\tfma.rn.f64 %3, %4, %8, %3;"
                 : "+d"(acc[4][4]), "+d"(acc[4][5]), "+d"(acc[4][6]), "+d"(acc[4][7])
                 : "d"(xa[4]), "d"(xs[4]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %3, %4, %0;\n\tfma.rn.f64 %1, %3, %5, %1;\n\tfma.rn.f64 %2, %3, %6, %2;"
                 : "+d"(acc[5][5]), "+d"(acc[5][6]), "+d"(acc[5][7])
                 : "d"(xa[5]), "d"(xs[5]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %2, %3, %0;\n\tfma.rn.f64 %1, %2, %4, %1;"
                 : "+d"(acc[6][6]), "+d"(acc[6][7])
                 : "d"(xa[6]), "d"(xs[6]), "d"(xs[7]));
    asm volatile("fma.rn.f64 %0, %1, %2, %0;"
                 : "+d"(acc[7][7])
                 : "d"(xa[7]), "d"(xs[7]));
  };
  auto load_row_diag = [&](uint32_t row_addr, double (&xa)[8]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) lds128(row_addr + boff_a[k], xa[2 * k], xa[2 * k + 1]);
  };
  // per-warp list of the tile's non-zero rows: {row byte offset in the stage, multiplicity as fp64},
  // built once per tile by all lanes, so that the row loop is a plain counted loop
  __shared__ __align__(16) double2 row_list[GRAM_WARPS][40];  // 32 rows + 8 zero-multiplicity pads
  double2* my_list = row_list[warp];
  const uint32_t list_addr = smem_u32(my_list);
  const uint32_t tiles_addr = smem_u32(tiles);
  const uint32_t row_bytes = (uint32_t)p.Ppad * 8u;
  auto load_entry = [&](int k, uint32_t& off, double& c) {
    double o;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(o), "=d"(c) : "r"(list_addr + 16u * (uint32_t)k));
    off = (uint32_t)__double_as_longlong(o);
  };
  // cross mode: per-warp scratch of (RT + 8) rows x Lpad scores behind the ring
  const int Lpad = p.ng * SLOT;
  double* my_scores = tiles + (size_t)p.stages * stage_doubles + (size_t)warp * (p.RT + 8) * Lpad;
  const uint32_t sc_addr = smem_u32(my_scores);
  const uint32_t sc_row_bytes = (uint32_t)Lpad * 8u;
  const double* wf_rep = CROSS ? p.wf + rep * p.Ppad : nullptr;
  if constexpr (CROSS) {
    for (int e = lane; e < (p.RT + 8) * Lpad; e += 32) my_scores[e] = 0.0;
    __syncwarp();
  }
  uint32_t cnt_next = load_counts(0);

  for (int t = 0; t < n_rt; ++t) {
    const int s = t % p.stages;
    const uint32_t use = (uint32_t)(t / p.stages);
    const uint32_t cnt = cnt_next;
    if (t + 1 < n_rt) cnt_next = load_counts(t + 1);  // prefetch: hides the global-load latency
    // rows of multiplicity 1 first, then the others
    const uint32_t mask1 = __ballot_sync(0xffffffffu, cnt == 1), mask2 = __ballot_sync(0xffffffffu, cnt > 1);
    const int n_one = __popc(mask1), n_nz = n_one + __popc(mask2);
    if (cnt != 0) {
      const uint32_t below = (1u << lane) - 1u;
      const int pos = (cnt == 1) ? __popc(mask1 & below) : n_one + __popc(mask2 & below);
      my_list[pos] = make_double2(__longlong_as_double((long long)((uint32_t)lane * row_bytes)), (double)cnt);
    }
    const int n_pairs_one = CROSS ? 0 : (n_one & ~1);  // an odd last multiplicity-1 row takes the scaled path
    // pads: multiplicity 0 on row 0 of the stage, so the pipelined loop below needs no branches
    if (lane < 8) my_list[n_nz + lane] = make_double2(__longlong_as_double(0ll), 0.0);
    __syncwarp();
    mbar_wait(&full_bar[s], use & 1);
    const uint32_t base = tiles_addr + (uint32_t)s * (uint32_t)(stage_doubles * 8);
    uint32_t release_dep = 0;
    if constexpr (CROSS) {
      // scores of the tile's non-zero rows for the LVs this warp's tiles touch, one (row, LV) pair
      // per lane:  scratch[k][l] = c_k * sum_{q in block l} x~[row_k][q] wf[q]   (pad rows: c = 0)
      if (n_nz > 0) {
        const int ns = p.Ppad / SLOT;
        const int lv_lo = ((tg * 32) / ns) * SLOT;
        const int lv_hi = min(p.L, (min(p.n_tiles - 1, tg * 32 + 31) / ns) * SLOT + SLOT);
        const int nlw = lv_hi - lv_lo;
        for (int e = lane; e < nlw * (n_nz + 2); e += 32) {
          const int k = e / nlw, lv = lv_lo + (e - k * nlw);
          const int slot0 = p.lv_off[lv] >> 3, nsl = (p.lv_k[lv] + SLOT - 1) >> 3;
          const int rot = (slot0 >> 1) & 3;
          uint32_t o;
          double cc, sc = 0.0;
          load_entry(k, o, cc);
          for (int sl = 0; sl < nsl; ++sl)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              const int col = (slot0 + sl) * SLOT + 2 * ((ch + rot) & 3);
              const double2 w2 = __ldg(reinterpret_cast<const double2*>(wf_rep + col));
              double x0, x1;
              lds128(base + o + (uint32_t)col * 8u, x0, x1);
              sc = fma(x0, w2.x, sc);
              sc = fma(x1, w2.y, sc);
            }
          my_scores[(size_t)k * Lpad + lv] = cc * sc;
        }
      }
      __syncwarp();
    }
    if (n_nz > 0 && diag) {
      // diagonal tile group: same software pipeline, one operand per row, 8 DMUL + 36 DFMA per row
      double xa0[8], xa1[8], xs[8], c0, c1, ce0, ce1;
      uint32_t oe0, oe1;
      load_entry(0, oe0, ce0);
      load_entry(1, oe1, ce1);
      load_row_diag(base + oe0, xa0);
      c0 = ce0;
      int k = 0;
      for (; k < n_pairs_one; k += 2) {
        load_row_diag(base + oe1, xa1);
        load_entry(k + 2, oe0, ce0);
        accumulate_diag(std::false_type{}, xa0, xs, 1.0);
        load_row_diag(base + oe0, xa0);
        load_entry(k + 3, oe1, ce1);
        accumulate_diag(std::false_type{}, xa1, xs, 1.0);
      }
      c0 = ce0;
      for (; k < n_nz; k += 2) {
        load_row_diag(base + oe1, xa1);
        c1 = ce1;
        load_entry(k + 2, oe0, ce0);
        accumulate_diag(std::true_type{}, xa0, xs, c0);
        load_row_diag(base + oe0, xa0);
        c0 = ce0;
        load_entry(k + 3, oe1, ce1);
        accumulate_diag(std::true_type{}, xa1, xs, c1);
      }
      asm volatile("{\n.reg .b32 lo, hi;\nmov.b64 {lo, hi}, %1;\nand.b32 %0, lo, 0;\n}" : "=r"(release_dep) : "d"(xa0[7]));
    } else if (n_nz > 0) {
      // Software pipeline over the non-zero rows, two rows per trip, straight-line body: the
      // operands of row k+1 are in flight (LDS) while the 64 FMAs of row k issue, and the list
      // entries of rows k+2 / k+3 are already in registers.  An odd row count runs one padded row
      // with multiplicity 0 (adds exact zeros; ~2.5 % extra FMAs, no branch in the body).
      double xa0[8], xb0[8], xa1[8], xb1[8], c0, c1, ce0, ce1;
      uint32_t oe0, oe1;
      // column operand: the same X row (Gram) or the row's scratch scores (cross)
      uint32_t bsrc = sc_addr;
      load_entry(0, oe0, ce0);
      load_entry(1, oe1, ce1);
      load_row(base + oe0, CROSS ? bsrc : base + oe0, xa0, xb0);
      c0 = ce0;
      int k = 0;
      for (; k < n_pairs_one; k += 2) {  // multiplicity 1: no scaling (Gram mode only)
        load_row(base + oe1, base + oe1, xa1, xb1);
        load_entry(k + 2, oe0, ce0);
        accumulate(std::false_type{}, xa0, xb0, 1.0);
        load_row(base + oe0, base + oe0, xa0, xb0);
        load_entry(k + 3, oe1, ce1);
        accumulate(std::false_type{}, xa1, xb1, 1.0);
      }
      c0 = ce0;
      for (; k < n_nz; k += 2) {
        load_row(base + oe1, CROSS ? bsrc + sc_row_bytes : base + oe1, xa1, xb1);
        c1 = ce1;
        load_entry(k + 2, oe0, ce0);
        accumulate(std::true_type{}, xa0, xb0, c0);
        bsrc += 2 * sc_row_bytes;
        load_row(base + oe0, CROSS ? bsrc : base + oe0, xa0, xb0);
        c0 = ce0;
        load_entry(k + 3, oe1, ce1);
        accumulate(std::true_type{}, xa1, xb1, c1);
      }
      // The loop prefetches one row set past the end (a pad row of this stage).  Make the stage release
      // below depend on that last load, so no shared-memory read of the stage is still in flight when
      // the refilling bulk copy may overwrite it.
      asm volatile("{\n.reg .b32 lo, hi;\nmov.b64 {lo, hi}, %1;\nand.b32 %0, lo, 0;\n}" : "=r"(release_dep) : "d"(xb0[7]));
    }
    __syncwarp();
    if (lane == 0) {
      // (release_dep == 0, but it makes this release depend on the warp's last shared-memory load)
      __threadfence_block();
      const int prev = atomicAdd(&stage_done[s] + release_dep, 1);
      if (prev == n_active - 1) {  // every consumer is done with this fill: refill the stage
        stage_done[s] = 0;
        if (t + p.stages < n_rt) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          issue_tile(t + p.stages);
        }
      }
    }
  }

  // ---- write the partial tile (undo the chunk rotation) and the column sums --------------------
  const size_t slab = (size_t)rep * p.n_chunks + chunk;
  if (tile_ok) {
    const int store_tile = CROSS ? sa * p.ng + sb : tile;
    double* g = p.G + (slab * p.n_tiles + store_tile) * TILE;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ra = 2 * (((i >> 1) + rot_a) & 3) + (i & 1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int cb = 2 * (((j >> 1) + rot_b) & 3) + (j & 1);
        g[ra * SLOT + cb] = (diag && j < i) ? acc[j][i] : acc[i][j];  // diagonal groups hold the upper triangle
      }
    }
  }
}

// Weighted column sums colsum[b][p] = sum_i c_bi x~_ip  (a skinny fp64 GEMM, counts x X~).
// CTA = 32 replicates x 256 columns over one row chunk; the X row tile and the (fp64-converted)
// multiplicities are staged in shared memory, thread = 8 replicates x 4 columns in registers
// (6 LDS.128 per 32 FMAs), two CTAs per SM overlap staging and arithmetic.
constexpr int CS_REPS = 32, CS_COLS = 256, CS_ROWS = 32;
__global__ void __launch_bounds__(256, 2) colsum_kernel(const double* __restrict__ X, const uint32_t* __restrict__ counts,
                                                        int64_t N, int Ppad, int64_t nrep, int n_chunks,
                                                        int64_t chunk_rows, double* __restrict__ out) {
  extern __shared__ __align__(16) double cs_smem[];
  double* xs = cs_smem;                       // [CS_ROWS][CS_COLS]
  double* cw = xs + CS_ROWS * CS_COLS;        // [CS_ROWS][CS_REPS]
  const int col0 = blockIdx.x * CS_COLS;
  const int64_t rep0 = (int64_t)blockIdx.y * CS_REPS;
  const int chunk = blockIdx.z;
  const int64_t r0 = (int64_t)chunk * chunk_rows, r1 = min(r0 + chunk_rows, N);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rg = warp >> 1;                   // replicate group: replicates 8*rg .. 8*rg+7
  const int cg = (warp & 1) * 32 + lane;      // column group: columns 4*cg .. 4*cg+3
  double acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int64_t row = r0; row < r1; row += CS_ROWS) {
    const int rows = (int)min((int64_t)CS_ROWS, r1 - row);
    __syncthreads();
    for (int e = threadIdx.x; e < CS_ROWS * CS_COLS; e += 256) {
      const int r = e / CS_COLS, c = e - r * CS_COLS;
      xs[e] = (r < rows && col0 + c < Ppad) ? X[(row + r) * Ppad + col0 + c] : 0.0;
    }
    for (int e = threadIdx.x; e < CS_ROWS * CS_REPS; e += 256) {
      const int b = e / CS_ROWS, r = e - b * CS_ROWS;  // consecutive threads: consecutive rows of one replicate
      double v = 0.0;
      if (r < rows && rep0 + b < nrep) v = counts ? (double)counts[(rep0 + b) * N + row + r] : 1.0;
      cw[r * CS_REPS + b] = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int r = 0; r < CS_ROWS; ++r) {
      const double2 x01 = *reinterpret_cast<const double2*>(&xs[r * CS_COLS + 4 * cg]);
      const double2 x23 = *reinterpret_cast<const double2*>(&xs[r * CS_COLS + 4 * cg + 2]);
      double c[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double2 c2 = *reinterpret_cast<const double2*>(&cw[r * CS_REPS + 8 * rg + 2 * k]);
        c[2 * k] = c2.x; c[2 * k + 1] = c2.y;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fma(c[i], x01.x, acc[i][0]); acc[i][1] = fma(c[i], x01.y, acc[i][1]);
        acc[i][2] = fma(c[i], x23.x, acc[i][2]); acc[i][3] = fma(c[i], x23.y, acc[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t b = rep0 + 8 * rg + i;
    if (b >= nrep) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + 4 * cg + j;
      if (c < Ppad) out[(b * n_chunks + chunk) * Ppad + c] = acc[i][j];
    }
  }
}

// sum of the per-chunk partials in chunk order (deterministic)
__global__ void reduce_chunks_kernel(const double* __restrict__ part, int64_t nrep, int n_chunks, int64_t per_rep,
                                     double* __restrict__ out) {
  const int64_t total = nrep * per_rep;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / per_rep, k = e - b * per_rep;
    const double* src = part + (b * n_chunks) * per_rep + k;
    // eight independent chains (eight loads in flight per thread: a single fit has few outputs and ~100 chunks,
    // one dependent chain made this kernel pure load latency), combined in a fixed order
    double a[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    int c = 0;
    for (; c + 8 <= n_chunks; c += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += src[(int64_t)(c + u) * per_rep];
    }
    for (int u = 0; c < n_chunks; ++c, ++u) a[u] += src[(int64_t)c * per_rep];
    out[e] = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
  }
}
