// Weighted second moments of a whole bootstrap batch on the tensor cores, exactly, without resident digit planes.
// Part of the single translation unit plspm_b200.cu (included there); see DESIGN.md §2 and §4.
//
//   G_b[p][q] = sum_i c_bi (x~_ip x~_iq),   colsum_b[p] = sum_i c_bi x~_ip        for every replicate b of the batch
//
// is the integer GEMM  counts [nb x N]  x  Z [N x pairs]  once Z is an integer matrix.  x~ is pre-scaled per column
// by a power of two to |x'| <= 2^23 and stored as h + l: h = rint(x') (int32), l = x' - h in [-1/2, 1/2] (fp32), i.e.
// 48 bits of x'.  A pair product is then
//       z = h_p h_q  (one IMAD.WIDE, exact)  +  rint( h_p l_q + l_p h_q + l_p l_q )  (three fp32 operations)
// and U = z + 2^47 is a 48-bit unsigned integer: six base-256 digits = the six low bytes of the register pair.
// (Column sums use x'_p * 2^23.)  The cross terms are rounded in fp32 (<= 1/4 unit each) before the final
// rounding, so |z - x'_p x'_q| <= ~1.25 units of 2^(e_p+e_q-46): ~2^-46 of the product of the column bounds -- 30
// times finer than the digit planes of round 1, random in sign.  Everything after it is exact integer arithmetic:
// u8 digits x s8 multiplicities, int32 accumulation (sum_i c_i <= N bounds every accumulator by 255 N < 2^31 for
// N < 8.4e6), int64 recombination.  No fp64 instruction runs in the hot loop: measured, the fp64 pipe delivers
// ~1/20 of its rate while the tensor pipe is busy with the int8 MMAs (in-kernel timers: 830 clk for 22 DFMA per row).
//
// One CTA = (one M tile of 128 digit rows, up to 512 replicates, one range of rows).  Per stage of 128 rows:
//   TMA        multiplicities c8 [512 x 128] (K-major, 128B swizzle), stored in HBM as the shared-memory image of the
//              tile (counts8_image_kernel), so a stage is ONE 64 KB linear bulk copy; three stages in the ring
//   generate   4 warps, thread = row: the 16 operands of its row (<= 2 column slots of x') come straight from L2
//              (coalesced: the copy is transposed, and they are prefetched one stage ahead), then the <= 23 fused
//              multiply-adds, the digits as a byte stream (pair-major, digit-minor: pair j owns stream bytes
//              6j..6j+5), stored as the MN-major A tile
//   MMA        tcgen05.mma kind::i8, M = 128, N <= 256 (two accumulators), K = 32 x 4, accumulators in TMEM
// Epilogue: the digits of a pair sit in 6 consecutive TMEM lanes; slabs of 32 replicates go through shared memory,
// are recombined to {lo = d0 + 2^8 d1 + 2^16 d2, hi = d3 + 2^8 d4 + 2^16 d5} (int64) and stored replicate-minor;
// gram_finalize_kernel adds row ranges and the two pieces of pairs that straddle M tiles, removes the offset
// 2^47 N and scales to fp64.  Tile kinds (what the 64 pair slots of a 384-byte stream are):
//   OFF(a,b)   x'_a[r] x'_b[c], r, c < 8            3 M tiles
//   DIAG(a)    x'_a[r] x'_a[c], r <= c  (36 pairs)  2 M tiles
//   SUM2(a,b)  x'_a[j] 2^23, x'_b[j] 2^23 (16)      1 M tile
#pragma once
#include "umma.cuh"

constexpr int GM_THREADS = 352, GM_STAGE_ROWS = 128, GM_PAIRS_PER_TILE = 22, GM_B_STAGES = 3;
constexpr uint32_t GM_B_BYTES = 512 * 128, GM_A_BYTES = 128 * 128;
enum { GM_KIND_OFF = 0, GM_KIND_DIAG = 1, GM_KIND_SUM2 = 2, GM_KIND_NONE = 3 /* padding tile of a cluster */ };

struct GramMmaParams {
  const int4* mtiles;   // [n_mtiles] {kind, third T, slot a, slot b}
  longlong2* part;      // [ksplit][n_mtiles * 22][nb_pad] {lo, hi}
  int64_t nb, nb_pad, N;
  int n_mtiles, n_groups, ksplit;
  int rows_per_cta;     // multiple of GM_STAGE_ROWS
  const uint8_t* c8img; // multiplicities of the batch as tile images: [stage of 128 rows][group of 512 replicates] 64 KB each
  unsigned long long* stats;  // optional [16]: pipeline diagnostics (PLSPM_KERNEL_STATS)
  const int2* XsT;      // [Ppad][ldx] pre-scaled transposed observations {h = rint(x'), bits of l = x' - h} (read straight from L2)
  int64_t ldx;
};

__host__ inline size_t gm_smem_bytes() { return 1024 + (size_t)GM_B_STAGES * GM_B_BYTES + (size_t)2 * GM_A_BYTES; }

// ---- which operands multiply in pair slot j of a tile kind (compile-time) ------------------------------------------
// returns a | b << 8 with a, b indices into the thread's 16 operand registers (slot a: 0..7, slot b: 8..15),
// b == 16: the constant 2^23 (column sums), a == 255: empty slot (zero digits)
__host__ __device__ constexpr int gm_pair_ops(int kind, int j) {
  if (kind == GM_KIND_OFF) return (j >> 3) | ((8 + (j & 7)) << 8);
  if (kind == GM_KIND_DIAG) {
    int r = 0, base = 0;
    while (r < 8 && j >= base + (8 - r)) { base += 8 - r; ++r; }
    return r < 8 ? (r | ((r + (j - base)) << 8)) : 255;
  }
  return j < 16 ? (j | (16 << 8)) : 255;
}

struct GmOperand { int h; float hf, l; };  // x' = h + l, hf = (float)h (exact: |h| <= 2^23)
template <int KIND, int J>
__device__ __forceinline__ void gm_digits(const GmOperand (&x)[16], uint32_t& lo, uint32_t& hi) {
  constexpr int ops = gm_pair_ops(KIND, J);
  if constexpr ((ops & 255) == 255) {
    lo = 0u; hi = 0u;
  } else {
    constexpr int ia = ops & 255, ib = ops >> 8;
    long long z;
    if constexpr (ib == 16) {  // x' * 2^23
      z = (long long)x[ia].h * 8388608ll + (long long)__float2int_rn(x[ia].l * 8388608.f);
    } else {
      const float cross = fmaf(x[ia].l, x[ib & 15].hf, fmaf(x[ib & 15].l, x[ia].hf, x[ia].l * x[ib & 15].l));
      z = (long long)x[ia].h * (long long)x[ib & 15].h + (long long)__float2int_rn(cross);
    }
    z += 140737488355328ll;  // 2^47
    lo = (uint32_t)z;
    hi = (uint32_t)(z >> 32);  // low 16 bits: digits 4, 5
  }
}

// Stream words [32 T, 32 T + 32) of a tile kind for one row.  Two consecutive pairs are 12 bytes = 3 words.
template <int KIND, int T, int CPL>
__device__ __forceinline__ void gm_couple(const GmOperand (&x)[16], uint32_t (&w)[32]) {
  constexpr int base = 3 * CPL - 32 * T;  // first of the couple's three words, relative to the tile
  if constexpr (base + 2 >= 0 && base < 32) {
    uint32_t lo_a, hi_a, lo_b, hi_b;
    gm_digits<KIND, 2 * CPL>(x, lo_a, hi_a);
    gm_digits<KIND, 2 * CPL + 1>(x, lo_b, hi_b);
    if constexpr (base >= 0 && base < 32) w[base] = lo_a;
    if constexpr (base + 1 >= 0 && base + 1 < 32) w[base + 1] = __byte_perm(hi_a, lo_b, 0x5410);
    if constexpr (base + 2 >= 0 && base + 2 < 32) w[base + 2] = __byte_perm(lo_b, hi_b, 0x5432);
  }
}
template <int KIND, int T, int... CPL>
__device__ __forceinline__ void gm_row_impl(const GmOperand (&x)[16], uint32_t (&w)[32], std::integer_sequence<int, CPL...>) {
  (gm_couple<KIND, T, CPL>(x, w), ...);
}
template <int KIND, int T>
__device__ __forceinline__ void gm_row(const GmOperand (&x)[16], uint32_t (&w)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) w[i] = 0u;
  gm_row_impl<KIND, T>(x, w, std::make_integer_sequence<int, 32>{});
}

__global__ void __launch_bounds__(GM_THREADS, 1) gram_mma_kernel(const GramMmaParams P) {
  using namespace umma;
  extern __shared__ uint8_t gm_smem_raw[];
  // barriers (8 bytes each): b_full[3] | b_empty[3] | a_full[2] | a_empty[2] | acc_full
  __shared__ uint64_t bars[11];
  __shared__ uint32_t tmem_base_sm;
  constexpr int B_BF = 0, B_BE = 3, B_AF = 6, B_AE = 8, B_ACC = 10;
  uint8_t* smem = gm_smem_raw + ((1024u - (s32(gm_smem_raw) & 1023u)) & 1023u);
  // shared memory: multiplicity ring 3 x 64 KB (two stages in flight while one is multiplied: a stage takes ~1 us to
  // arrive when the whole chip is pulling), then the digit (A) ring 2 x 16 KB
  uint8_t* aring = smem + GM_B_STAGES * GM_B_BYTES;

  int t = blockIdx.x;
  const int mt = t % P.n_mtiles; t /= P.n_mtiles;
  const int grp = t % P.n_groups;
  const int ks = t / P.n_groups;
  const int4 desc = P.mtiles[mt];
  const int kind = desc.x, T = desc.y, slot_a = desc.z, slot_b = desc.w;
  const int64_t b0 = (int64_t)grp * 512;
  const int ncols = (int)min((int64_t)512, P.nb - b0);
  const int n0 = min(256, (ncols + 15) & ~15), n1 = ncols > 256 ? ((ncols - 256 + 15) & ~15) : 0;
  const int64_t row_begin = (int64_t)ks * P.rows_per_cta;  // (a multiple of GM_STAGE_ROWS)
  const int64_t row_end = min(P.N, row_begin + P.rows_per_cta);
  const int n_stages = row_end > row_begin ? (int)((row_end - row_begin + GM_STAGE_ROWS - 1) / GM_STAGE_ROWS) : 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) tmem_alloc(&tmem_base_sm, 512);
  if (threadIdx.x == 0) {
    for (int s = 0; s < GM_B_STAGES; ++s) { bar_init(&bars[B_BF + s], 1); bar_init(&bars[B_BE + s], 2); }
    for (int s = 0; s < 2; ++s) { bar_init(&bars[B_AF + s], 4); bar_init(&bars[B_AE + s], 2); }
    bar_init(&bars[B_ACC], 2);
    bar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_sm;
  const uint32_t bar0 = s32(bars), smem0 = s32(smem), aring0 = s32(aring);

  if (warp == 0) {
    if (lane == 0) {  // ---- producer: one linear bulk copy per stage (the multiplicities are stored as tile images) -----
      const uint32_t bytes = n1 ? GM_B_BYTES : GM_B_BYTES / 2;
      const uint8_t* src = P.c8img + ((size_t)(row_begin / GM_STAGE_ROWS) * P.n_groups + grp) * GM_B_BYTES;
      const size_t step = (size_t)P.n_groups * GM_B_BYTES;
      int s = 0, ph = 1;
      long long w11 = 0;
      for (int j = 0; j < n_stages; ++j, src += step) {
        { const long long c_ = P.stats ? clock64() : 0; bar_wait_a(bar0 + 8 * (B_BE + s), ph, 11); if (P.stats) w11 += clock64() - c_; }
        bar_expect_tx_a(bar0 + 8 * (B_BF + s), bytes);
        bulk_load_a(smem0 + (uint32_t)s * GM_B_BYTES, src, bytes, bar0 + 8 * (B_BF + s));
        if (++s == GM_B_STAGES) { s = 0; ph ^= 1; }
      }
      if (P.stats) atomicAdd(P.stats + 1, (unsigned long long)w11);
    }
  } else if (warp == 1 || warp == 6) {
    // ---- two MMA issuers, one per accumulator (replicates 0..255 / 256..511): a single issuing thread needs ~1000 clk
    // per stage for its 8 MMAs, 2 waits and 2 commits -- as long as the tensor core needs for the stage
    if (lane == 0 && n_stages > 0) {
      const int acc = warp == 1 ? 0 : 1;
      const int nn = acc == 0 ? n0 : n1;
      const uint32_t idesc = instr_desc(D_S32, AB_U8, AB_S8, 1, 0, 128, (uint32_t)(nn ? nn : 16));
      const uint64_t adesc0 = smem_desc(aring0, 16, 1024, SW_128B);                                  // MN-major digits
      const uint64_t bdesc0 = smem_desc(smem0 + acc * (GM_B_BYTES / 2), 16, 1024, SW_128B);          // K-major multiplicities
      const uint32_t dcol = tbase + 256 * acc;
      long long w12 = 0, w13 = 0;
      const long long t_begin = clock64();
      int s = 0, ph = 0;
      for (int j = 0; j < n_stages; ++j) {
        const int a = j & 1;
        { const long long c_ = P.stats ? clock64() : 0; bar_wait_a(bar0 + 8 * (B_BF + s), ph, 12); if (P.stats) w12 += clock64() - c_; }
        { const long long c_ = P.stats ? clock64() : 0; bar_wait_a(bar0 + 8 * (B_AF + a), (j >> 1) & 1, 13); if (P.stats) w13 += clock64() - c_; }
        tc_fence_after();
        if (nn) {
          const uint64_t adesc = adesc0 + (uint64_t)(a * (GM_A_BYTES >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)(s * (GM_B_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < GM_STAGE_ROWS / 32; ++k) mma_i8_ss(dcol, adesc + 256 * k, bd + 2 * k, idesc, (j | k) ? 1u : 0u);
        }
        // (both issuers release: the stage and the digit buffer are free when the MMAs of both accumulators are done)
        mma_commit_a(bar0 + 8 * (B_BE + s));
        mma_commit_a(bar0 + 8 * (B_AE + a));
        if (++s == GM_B_STAGES) { s = 0; ph ^= 1; }
      }
      mma_commit_a(bar0 + 8 * B_ACC);
      if (P.stats && acc == 0) {
        atomicAdd(P.stats + 2, (unsigned long long)w12);
        atomicAdd(P.stats + 3, (unsigned long long)w13);
        atomicAdd(P.stats + 0, (unsigned long long)(clock64() - t_begin));
        atomicAdd(P.stats + 9, (unsigned long long)n_stages);
      }
    }
  } else {
    // ---- 2 x 4 generator / epilogue warps: group 0 = warps 2..5 takes the even stages (digit buffer 0), group 1 =
    // warps 7..10 the odd ones (digit buffer 1).  Generating a stage is a ~700 clk chain for one warp per scheduler;
    // two stages in flight keep it off the critical path.  (TMEM lane quadrant = warp % 4: both groups cover all four.)
    const int grp = warp >= 7 ? 1 : 0;
    const int g = threadIdx.x - (grp ? 224 : 64);  // 0..127: row of the stage
    // operands of this thread's row: column slots a and b of x' = {h, l} (transposed copy: the 32 lanes read 256
    // contiguous bytes), prefetched one of the group's stages (= two stages) ahead: every SM streams its multiplicity
    // tiles at the same time and the loads take longer than a stage
    const int2* xa = P.XsT + (size_t)slot_a * 8 * P.ldx;
    const int2* xb = P.XsT + (size_t)slot_b * 8 * P.ldx;
    int2 xn[16];
    auto fetch = [&](int j) {
      const int64_t i = row_begin + (int64_t)j * GM_STAGE_ROWS + g;
      const bool ok = j < n_stages && i < P.N;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        xn[c] = ok ? __ldg(xa + (size_t)c * P.ldx + i) : make_int2(0, 0);
        xn[8 + c] = ok ? __ldg(xb + (size_t)c * P.ldx + i) : make_int2(0, 0);
      }
    };
    fetch(grp);
    const uint32_t arow0 = (uint32_t)g * 128, gsw = (uint32_t)(g & 7);
    long long w14 = 0, wgen = 0, wld = 0;
    const long long tg_begin = clock64();
    for (int j = grp; j < n_stages; j += 2) {
      const long long cg_ = P.stats ? clock64() : 0;
      const int a = grp;
      GmOperand x[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) { x[c].h = xn[c].x; x[c].hf = (float)xn[c].x; x[c].l = __int_as_float(xn[c].y); }
      if (P.stats) {  // (diagnostics: how long the prefetched operands take to arrive)
        int sx = 0;
#pragma unroll
        for (int c = 0; c < 16; ++c) sx += x[c].h;
        if (sx == 0x7fffffff) wgen = 1;
        wld += clock64() - cg_;
      }
      fetch(j + 2);
      uint32_t w[32];
      switch (kind * 4 + T) {
        case GM_KIND_OFF * 4 + 0: gm_row<GM_KIND_OFF, 0>(x, w); break;
        case GM_KIND_OFF * 4 + 1: gm_row<GM_KIND_OFF, 1>(x, w); break;
        case GM_KIND_OFF * 4 + 2: gm_row<GM_KIND_OFF, 2>(x, w); break;
        case GM_KIND_DIAG * 4 + 0: gm_row<GM_KIND_DIAG, 0>(x, w); break;
        case GM_KIND_DIAG * 4 + 1: gm_row<GM_KIND_DIAG, 1>(x, w); break;
        case GM_KIND_SUM2 * 4 + 0: gm_row<GM_KIND_SUM2, 0>(x, w); break;
        default:
#pragma unroll
          for (int i = 0; i < 32; ++i) w[i] = 0u;
          break;
      }
      if (P.stats) wgen += clock64() - cg_;
      // the digits are ready in registers; the A buffer is free once the MMAs of stage j - 2 have read it
      { const long long c_ = P.stats ? clock64() : 0; bar_wait_a(bar0 + 8 * (B_AE + a), ((j >> 1) & 1) ^ 1, 14); if (P.stats) w14 += clock64() - c_; }
      // MN-major A tile, 128-byte swizzle: byte (m, k) at k * 128 + ((m / 16) ^ (k % 8)) * 16 + m % 16
      uint8_t* arow = aring + (size_t)a * GM_A_BYTES + arow0;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(arow + ((q ^ gsw) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) bar_arrive_a(bar0 + 8 * (B_AF + a));
    }
    if (P.stats && g == 0 && grp == 0) {
      atomicAdd(P.stats + 4, (unsigned long long)w14);
      atomicAdd(P.stats + 5, (unsigned long long)wgen);
      atomicAdd(P.stats + 8, (unsigned long long)wld);
      atomicAdd(P.stats + 6, (unsigned long long)(clock64() - tg_begin));
    }
    const long long te_begin = clock64();
    if (n_stages > 0) {
      // ---- epilogue: slabs of 32 replicates through shared memory, digits -> {lo, hi} ----------------------------
      bar_wait_a(bar0 + 8 * B_ACC, 0, 15);
      tc_fence_after();
      int32_t* slab = reinterpret_cast<int32_t*>(smem) + grp * (128 * 33);  // [128 lanes][33] per group (all stages are drained)
      const int q = warp & 3;
      const int m = 32 * q + lane;  // TMEM lane = digit row of the tile
      longlong2* out = P.part + ((size_t)ks * P.n_mtiles + mt) * GM_PAIRS_PER_TILE * P.nb_pad;
      for (int c0 = 32 * grp; c0 < ncols; c0 += 64) {  // the groups take the 32-replicate slabs in turn
        uint32_t v[32];
        tmem_ld32(tbase + ((uint32_t)(32 * q) << 16) + c0, v);
        tmem_wait_ld();
        if (grp) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 1, 128;" ::: "memory");  // previous slab consumed
#pragma unroll
        for (int c = 0; c < 32; ++c) slab[m * 33 + c] = (int32_t)v[c];
        if (grp) asm volatile("bar.sync 2, 128;" ::: "memory"); else asm volatile("bar.sync 1, 128;" ::: "memory");
        for (int e = g; e < GM_PAIRS_PER_TILE * 32; e += 128) {
          const int lp = e >> 5, c = e & 31;
          if (c0 + c >= ncols) continue;
          long long lo = 0, hi = 0;
#pragma unroll
          for (int d = 0; d < 6; ++d) {
            const int mm = 6 * lp + d - 2 * T;  // lane of digit d of local pair lp
            const long long sv = (mm >= 0 && mm < 128) ? (long long)slab[mm * 33 + c] : 0ll;
            if (d < 3) lo += sv << (8 * d);
            else hi += sv << (8 * (d - 3));
          }
          out[(size_t)lp * P.nb_pad + b0 + c0 + c] = make_longlong2(lo, hi);
        }
      }
    }
    if (P.stats && g == 0 && grp == 0) atomicAdd(P.stats + 7, (unsigned long long)(clock64() - te_begin));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tbase, 512);
}

// One output moment = the sum over row ranges (and over the two M tiles a straddling pair lives in) of {lo, hi}:
//   value = (lo + 2^24 (hi - 2^23 N)) * dscale_p * dscale_q        (the offset 2^47 per row, times sum_i c_i = N)
struct GramOut {
  int slot0, slot1;  // piece slots (tile * 22 + local pair); slot1 = -1 when the pair lives in one tile
  int p, q;          // padded columns (q = -1: column sum of p)
  int dst1, dst2;    // offsets into the replicate's tile array (dst2: mirrored entry of a diagonal tile, or -1);
                     // column sums: dst1 = column
};
__global__ void __launch_bounds__(256) gram_finalize_kernel(const longlong2* __restrict__ part, int64_t nb, int64_t nb_pad,
                                                            int n_slots, int ksplit, const GramOut* __restrict__ outs,
                                                            int n_outs, const double* __restrict__ xunit, double N,
                                                            int64_t g_stride, double* __restrict__ G, int64_t cs_stride,
                                                            double* __restrict__ colsum) {
  const int64_t total = (int64_t)n_outs * nb;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(e / nb);
    const int64_t b = e - (int64_t)o * nb;
    const GramOut go = outs[o];
    long long lo = 0, hi = 0;
    for (int k = 0; k < ksplit; ++k) {
      const longlong2 a = part[((size_t)k * n_slots + go.slot0) * nb_pad + b];
      lo += a.x; hi += a.y;
      if (go.slot1 >= 0) {
        const longlong2 c = part[((size_t)k * n_slots + go.slot1) * nb_pad + b];
        lo += c.x; hi += c.y;
      }
    }
    const double unit = xunit[go.p] * (go.q >= 0 ? xunit[go.q] : 1.1920928955078125e-07);  // 2^-23
    const double v = ((double)lo + 16777216.0 * ((double)hi - 8388608.0 * N)) * unit;
    if (go.q < 0) {
      colsum[b * cs_stride + go.dst1] = v;
    } else {
      double* g = G + b * g_stride;
      g[go.dst1] = v;
      if (go.dst2 >= 0) g[go.dst2] = v;
    }
  }
}

// The multiplicities of a batch as shared-memory tile images (what the producers of gram_mma_kernel and
// vote_mma_kernel fetch with one linear bulk copy): image (stage S of 128 rows, group G of 512 replicates) =
// 64 KB at ((S * n_groups + G) * 64 KB) = two K-major [256 replicates x 128 B] tiles with the 128-byte swizzle,
// byte (replicate r, row k) at (r >> 8) * 32 KB + (r & 255) * 128 + (((k >> 4) ^ (r & 7)) << 4) + (k & 15).
// Replicates >= nb and rows >= N are zero.  Multiplicities above 127 do not fit the s8 operand: flag (fp64 fallback).
__global__ void __launch_bounds__(256) counts8_image_kernel(const uint32_t* __restrict__ counts, int64_t N, int64_t nb, int n_groups,
                                                            int64_t n_stages, uint8_t* __restrict__ img, int* __restrict__ overflow) {
  const int64_t total = n_stages * n_groups * 512 * 8;  // 16-byte chunks
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e & 7), r = (int)((e >> 3) & 511);
    const int64_t sg = e >> 12, S = sg / n_groups;
    const int64_t b = (sg - S * n_groups) * 512 + r, i0 = S * 128 + 16 * c;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (b < nb && i0 < N) {
      const uint32_t* src = counts + b * N + i0;
      bool big = false;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const uint32_t v = (i0 + k < N) ? src[k] : 0u;
        big |= v > 127u;
        w[k >> 2] |= (v & 127u) << (8 * (k & 3));
      }
      if (big) *overflow = 1;
    }
    *reinterpret_cast<uint4*>(img + (size_t)sg * 65536 + (size_t)(r >> 8) * 32768 + (size_t)(r & 255) * 128 + ((c ^ (r & 7)) << 4)) =
        make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// x' = x~ 2^(23 - e_p) (exact), transposed and split: XsT[p][i] = {rint(x'), fp32(x' - rint(x'))}; xunit[p] = 2^(e_p - 23)
// is the value of one unit of x'_p
__global__ void gram_xunit_kernel(const double* __restrict__ absmax_partial, int nblocks, int Ppad, double* __restrict__ xunit,
                                  double* __restrict__ xscale) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per column (warp_column_reduce)
  if (p >= Ppad) return;
  const double m = warp_column_reduce(absmax_partial, nblocks, (int64_t)Ppad, p, 0.0, OpMax());
  if (threadIdx.x & 31) return;
  int e = 0;
  if (m > 0.0) frexp(m, &e);  // m = f 2^e, f in [0.5, 1): |x~| < 2^e
  xunit[p] = ldexp(1.0, e - 23);
  xscale[p] = ldexp(1.0, 23 - e);
}
// Heavy-tail guard: rows of column p whose magnitude is within 2^-11 of the column bound.  Products are rounded to
// 2^-47 of the product of the column BOUNDS; if a handful of gross outliers set the bound (<= 8 such rows), the bulk of
// the column sits > 11 bits below it and a replicate that misses the outliers would see its variance at < 1e-8
// relative precision only up to there -- such data takes the fp64 kernels instead (see plspm_data_create).
__global__ void gram_tailcount_partial_kernel(const double* __restrict__ X, int64_t N, int Ppad, int64_t rows_per_block,
                                              const double* __restrict__ xunit, int* __restrict__ partial) {
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, N);
  for (int p = threadIdx.x; p < Ppad; p += blockDim.x) {
    const double thr = xunit[p] * 4096.0;  // 2^(e_p - 11)
    int n = 0;
    for (int64_t i = r0; i < r1; ++i) n += fabs(X[i * Ppad + p]) >= thr ? 1 : 0;
    partial[(int64_t)blockIdx.x * Ppad + p] = n;
  }
}
__global__ void gram_tailcount_final_kernel(const int* __restrict__ partial, int nblocks, int Ppad, int* __restrict__ count) {
  const int p = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // one warp per column
  if (p >= Ppad) return;
  const int n = warp_column_reduce(partial, nblocks, (int64_t)Ppad, p, 0, OpAdd());
  if ((threadIdx.x & 31) == 0) count[p] = n;
}
__global__ void __launch_bounds__(256) gram_xst_kernel(const double* __restrict__ X, int64_t N, int Ppad, int64_t ldx,
                                                       const double* __restrict__ xscale, int2* __restrict__ XsT) {
  __shared__ double tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int p0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t i = i0 + rr;
    const int p = p0 + tx;
    tile[rr][tx] = (i < N && p < Ppad) ? X[i * Ppad + p] * xscale[p] : 0.0;
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int p = p0 + rr;
    const int64_t i = i0 + tx;
    if (p < Ppad && i < ldx) {
      const double v = tile[tx][rr], h = rint(v);
      XsT[(int64_t)p * ldx + i] = make_int2((int)h, __float_as_int((float)(v - h)));
    }
  }
}
