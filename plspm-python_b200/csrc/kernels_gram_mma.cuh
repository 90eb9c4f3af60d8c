// Weighted second moments of a whole bootstrap batch on the tensor cores, exactly, without resident digit planes.
// Part of the single translation unit plspm_b200.cu (included there); see DESIGN.md §2 and §4.
//
//   G_b[p][q] = sum_i c_bi (x~_ip x~_iq),   colsum_b[p] = sum_i c_bi x~_ip        for every replicate b of the batch
//
// is the integer GEMM  counts [nb x N]  x  Z [N x pairs]  once Z is an integer matrix.  x~ is pre-scaled per column
// by a power of two to |x'| <= 2^23 (exact), so z = x'_p x'_q < 2^46, and ONE fused multiply-add
//       r = fma(x'_p, x'_q, 2^52 + 2^51 + 2^47)
// rounds z to the nearest integer q and leaves U = q + 2^47 in the low 48 mantissa bits of r: six unsigned base-256
// digits, i.e. the six low bytes of r's register pair.  (Column sums use x'_p * 2^23.)  Rounding: half a unit of
// 2^(e_p+e_q-46), i.e. 2^-47 of the product of the column bounds -- 64 times finer than the digit planes of round 1,
// random in sign; everything after it is exact integer arithmetic: u8 digits x s8 multiplicities, int32 accumulation
// (sum_i c_i <= N bounds every accumulator by 255 N < 2^31 for N < 8.4e6), int64 recombination.
//
// One CTA = (one M tile of 128 digit rows, up to 512 replicates, one range of rows).  Per stage of 128 rows:
//   TMA        multiplicities c8 [512 x 128] (K-major, 128B swizzle) and the <= 2 column slots of x' the tile needs
//              ([8 columns x 128 rows] fp64 each, from the transposed pre-scaled copy)
//   generate   4 warps, thread = row: the <= 23 fused multiply-adds of its row, the digits as a byte stream
//              (pair-major, digit-minor: pair j owns stream bytes 6j..6j+5), stored as the MN-major A tile
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

constexpr int GM_THREADS = 192, GM_STAGE_ROWS = 128, GM_STAGES = 2, GM_PAIRS_PER_TILE = 22;
constexpr uint32_t GM_B_BYTES = 512 * 128, GM_X_BYTES = 2 * 8 * 128 * 8, GM_A_BYTES = 128 * 128;
constexpr uint32_t GM_STAGE_BYTES = GM_B_BYTES + GM_X_BYTES + GM_A_BYTES;
enum { GM_KIND_OFF = 0, GM_KIND_DIAG = 1, GM_KIND_SUM2 = 2 };

struct GramMmaParams {
  const int4* mtiles;   // [n_mtiles] {kind, third T, slot a, slot b}
  longlong2* part;      // [ksplit][n_mtiles * 22][nb_pad] {lo, hi}
  int64_t nb, nb_pad, N;
  int n_mtiles, n_groups, ksplit;
  int rows_per_cta;     // multiple of GM_STAGE_ROWS
};

__host__ inline size_t gm_smem_bytes() { return 1024 + (size_t)GM_STAGES * GM_STAGE_BYTES; }

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

template <int KIND, int J>
__device__ __forceinline__ void gm_digits(const double (&x)[16], uint32_t& lo, uint32_t& hi) {
  constexpr int ops = gm_pair_ops(KIND, J);
  if constexpr ((ops & 255) == 255) {
    lo = 0u; hi = 0u;
  } else {
    constexpr int ia = ops & 255, ib = ops >> 8;
    const double r = fma(x[ia], ib == 16 ? 8388608.0 : x[ib & 15], 6896136929411072.0);  // 2^52 + 2^51 + 2^47
    lo = (uint32_t)__double2loint(r);
    hi = (uint32_t)__double2hiint(r);  // low 16 bits: digits 4, 5
  }
}

// Stream words [32 T, 32 T + 32) of a tile kind for one row.  Two consecutive pairs are 12 bytes = 3 words.
template <int KIND, int T, int CPL>
__device__ __forceinline__ void gm_couple(const double (&x)[16], uint32_t (&w)[32]) {
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
__device__ __forceinline__ void gm_row_impl(const double (&x)[16], uint32_t (&w)[32], std::integer_sequence<int, CPL...>) {
  (gm_couple<KIND, T, CPL>(x, w), ...);
}
template <int KIND, int T>
__device__ __forceinline__ void gm_row(const double (&x)[16], uint32_t (&w)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) w[i] = 0u;
  gm_row_impl<KIND, T>(x, w, std::make_integer_sequence<int, 32>{});
}

__global__ void __launch_bounds__(GM_THREADS, 1)
    gram_mma_kernel(const __grid_constant__ CUtensorMap map_c8, const __grid_constant__ CUtensorMap map_xs, const GramMmaParams P) {
  using namespace umma;
  extern __shared__ uint8_t gm_smem_raw[];
  __shared__ uint64_t in_full[GM_STAGES], a_full[GM_STAGES], empty[GM_STAGES], acc_full;
  __shared__ uint32_t tmem_base_sm;
  uint8_t* smem = gm_smem_raw + ((1024u - (s32(gm_smem_raw) & 1023u)) & 1023u);

  int t = blockIdx.x;
  const int mt = t % P.n_mtiles; t /= P.n_mtiles;
  const int grp = t % P.n_groups;
  const int ks = t / P.n_groups;
  const int4 desc = P.mtiles[mt];
  const int kind = desc.x, T = desc.y, slot_a = desc.z, slot_b = desc.w;
  const int64_t b0 = (int64_t)grp * 512;
  const int ncols = (int)min((int64_t)512, P.nb - b0);
  const int n0 = min(256, (ncols + 15) & ~15), n1 = ncols > 256 ? ((ncols - 256 + 15) & ~15) : 0;
  const int64_t row_begin = (int64_t)ks * P.rows_per_cta;
  const int64_t row_end = min(P.N, row_begin + P.rows_per_cta);
  const int n_stages = row_end > row_begin ? (int)((row_end - row_begin + GM_STAGE_ROWS - 1) / GM_STAGE_ROWS) : 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 1) tmem_alloc(&tmem_base_sm, 512);
  if (threadIdx.x == 0) {
    for (int s = 0; s < GM_STAGES; ++s) { bar_init(&in_full[s], 1); bar_init(&a_full[s], 4); bar_init(&empty[s], 1); }
    bar_init(&acc_full, 1);
    bar_fence_init();
    tma_prefetch_desc(&map_c8);
    tma_prefetch_desc(&map_xs);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_sm;

  if (warp == 0) {
    if (lane == 0) {  // ---- TMA producer -------------------------------------------------------------------------
      const uint32_t tx = (n1 ? GM_B_BYTES : GM_B_BYTES / 2) + (kind == GM_KIND_DIAG ? GM_X_BYTES / 2 : GM_X_BYTES);
      for (int j = 0; j < n_stages; ++j) {
        const int s = j % GM_STAGES;
        bar_wait(&empty[s], ((j / GM_STAGES) & 1) ^ 1, 11);
        uint8_t* st = smem + (size_t)s * GM_STAGE_BYTES;
        const int i0 = (int)(row_begin + (int64_t)j * GM_STAGE_ROWS);
        bar_expect_tx(&in_full[s], tx);
        tma_load_2d(st, &map_c8, &in_full[s], i0, (int)b0);
        if (n1) tma_load_2d(st + GM_B_BYTES / 2, &map_c8, &in_full[s], i0, (int)b0 + 256);
        tma_load_2d(st + GM_B_BYTES, &map_xs, &in_full[s], i0, slot_a * 8);
        if (kind != GM_KIND_DIAG) tma_load_2d(st + GM_B_BYTES + GM_X_BYTES / 2, &map_xs, &in_full[s], i0, slot_b * 8);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && n_stages > 0) {  // ---- MMA issuer -------------------------------------------------------------
      const uint32_t idesc0 = instr_desc(D_S32, AB_U8, AB_S8, 1, 0, 128, (uint32_t)n0);
      const uint32_t idesc1 = instr_desc(D_S32, AB_U8, AB_S8, 1, 0, 128, (uint32_t)(n1 ? n1 : 16));
      for (int j = 0; j < n_stages; ++j) {
        const int s = j % GM_STAGES;
        bar_wait(&in_full[s], (j / GM_STAGES) & 1, 12);
        bar_wait(&a_full[s], (j / GM_STAGES) & 1, 13);
        tc_fence_after();
        uint8_t* st = smem + (size_t)s * GM_STAGE_BYTES;
        const uint64_t adesc = smem_desc(s32(st + GM_B_BYTES + GM_X_BYTES), 16, 1024, SW_128B);  // MN-major
        const uint64_t bdesc0 = smem_desc(s32(st), 16, 1024, SW_128B);                           // K-major
        const uint64_t bdesc1 = smem_desc(s32(st + GM_B_BYTES / 2), 16, 1024, SW_128B);
#pragma unroll
        for (int k = 0; k < GM_STAGE_ROWS / 32; ++k) {
          mma_i8_ss(tbase, desc_advance(adesc, 4096 * k), desc_advance(bdesc0, 32 * k), idesc0, (j | k) ? 1u : 0u);
          if (n1) mma_i8_ss(tbase + 256, desc_advance(adesc, 4096 * k), desc_advance(bdesc1, 32 * k), idesc1, (j | k) ? 1u : 0u);
        }
        mma_commit(&empty[s]);
      }
      mma_commit(&acc_full);
    }
  } else {  // ---- 4 generator / epilogue warps (TMEM lane quadrant = warp % 4) ---------------------------------------
    const int g = threadIdx.x - 64;  // 0..127: row of the stage
    for (int j = 0; j < n_stages; ++j) {
      const int s = j % GM_STAGES;
      bar_wait(&in_full[s], (j / GM_STAGES) & 1, 14);  // (the A buffer of the stage is free: its MMAs released the stage)
      uint8_t* st = smem + (size_t)s * GM_STAGE_BYTES;
      const double* xs = reinterpret_cast<const double*>(st + GM_B_BYTES);
      double x[16];
#pragma unroll
      for (int c = 0; c < 8; ++c) x[c] = xs[c * 128 + g];
      if (kind != GM_KIND_DIAG) {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[8 + c] = xs[(8 + c) * 128 + g];
      } else {
#pragma unroll
        for (int c = 0; c < 8; ++c) x[8 + c] = 0.0;
      }
      uint32_t w[32];
      switch (kind * 4 + T) {
        case GM_KIND_OFF * 4 + 0: gm_row<GM_KIND_OFF, 0>(x, w); break;
        case GM_KIND_OFF * 4 + 1: gm_row<GM_KIND_OFF, 1>(x, w); break;
        case GM_KIND_OFF * 4 + 2: gm_row<GM_KIND_OFF, 2>(x, w); break;
        case GM_KIND_DIAG * 4 + 0: gm_row<GM_KIND_DIAG, 0>(x, w); break;
        case GM_KIND_DIAG * 4 + 1: gm_row<GM_KIND_DIAG, 1>(x, w); break;
        default: gm_row<GM_KIND_SUM2, 0>(x, w); break;
      }
      // MN-major A tile, 128-byte swizzle: byte (m, k) at k * 128 + ((m / 16) ^ (k % 8)) * 16 + m % 16
      uint8_t* arow = st + GM_B_BYTES + GM_X_BYTES + (size_t)g * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q)
        *reinterpret_cast<uint4*>(arow + ((q ^ (g & 7)) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) bar_arrive(&a_full[s]);
    }
    if (n_stages > 0) {
      // ---- epilogue: slabs of 32 replicates through shared memory, digits -> {lo, hi} ----------------------------
      bar_wait(&acc_full, 0, 15);
      tc_fence_after();
      int32_t* slab = reinterpret_cast<int32_t*>(smem);  // [128 lanes][33]  (all stages are drained)
      const int q = warp & 3;
      const int m = 32 * q + lane;  // TMEM lane = digit row of the tile
      longlong2* out = P.part + ((size_t)ks * P.n_mtiles + mt) * GM_PAIRS_PER_TILE * P.nb_pad;
      for (int c0 = 0; c0 < ncols; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(tbase + ((uint32_t)(32 * q) << 16) + (c0 < 256 ? c0 : 256 + (c0 - 256)), v);
        tmem_wait_ld();
        asm volatile("bar.sync 1, 128;" ::: "memory");  // previous slab consumed
#pragma unroll
        for (int c = 0; c < 32; ++c) slab[m * 33 + c] = (int32_t)v[c];
        asm volatile("bar.sync 1, 128;" ::: "memory");
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

// x' = x~ 2^(23 - e_p) (exact), transposed: XsT[p][i]; xunit[p] = 2^(e_p - 23) is the value of one unit of x'_p
__global__ void gram_xunit_kernel(const double* __restrict__ absmax_partial, int nblocks, int Ppad, double* __restrict__ xunit,
                                  double* __restrict__ xscale) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ppad) return;
  double m = 0.0;
  for (int b = 0; b < nblocks; ++b) m = fmax(m, absmax_partial[(int64_t)b * Ppad + p]);
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
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= Ppad) return;
  int n = 0;
  for (int b = 0; b < nblocks; ++b) n += partial[(int64_t)b * Ppad + p];
  count[p] = n;
}
__global__ void __launch_bounds__(256) gram_xst_kernel(const double* __restrict__ X, int64_t N, int Ppad, int64_t ldx,
                                                       const double* __restrict__ xscale, double* __restrict__ XsT) {
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
    if (p < Ppad && i < ldx) XsT[(int64_t)p * ldx + i] = tile[tx][rr];
  }
}
