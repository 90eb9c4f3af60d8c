// Fused tensor-core sign vote (tcgen05 + TMEM + TMA): replaces score generation + a library GEMM.
// Part of the single translation unit plspm_b200.cu (included there); see DESIGN.md §4.
//
//   E'[p][l][b] = sum_i xh_ip * fp16( c_bi * t'_bil ),     t'_bil = sum_{q in block l} xh_iq w'_bq,  w' = wf * sd
//
// (xh = x~/sd in fp16; t' is the UN-centred score: the solver removes sh_l * sum_i c_bi xh_ip afterwards, see
// solver_core.h phase 3.)  Only the SIGN of E is used, and only where it exceeds a rigorous error bound.
//
// One CTA = (128 replicates, one latent variable l, <= 256 manifest columns p, one range of rows).  All three operands
// live in HBM as ready-made shared-memory tile images (vote_xt_image_kernel, vote_xl_image_kernel,
// counts8_image_kernel), so a ring stage (= one chunk of 64 rows, four stages) is three linear bulk copies:
//   XhT image   [256 p x 64 k] fp16, K-major, 128B swizzle (32 KB): B operand of the vote MMA
//   c8 image    [128 replicates x 64 rows] multiplicities, 64B swizzle (8 KB)
//   Xh_blk      [64 rows x 16 columns] fp16 per K = 16 step of the block, 32B swizzle (2 KB each): B of the score MMA
// Per chunk of 64 rows, everything stays on the SM:
//   MMA 1     D1[b][i] = W_l[b][:] . Xh_blk[i][:]       (M = 128 replicates, N = 64 rows, K = 16 per step, SS)
//             -> the scores of 128 replicates x 64 rows, fp32, in tensor memory
//   epilogue  tcgen05.ld the scores, multiply by the row multiplicity, round to fp16, tcgen05.st them back to tensor
//             memory as the A operand of
//   MMA 2     E[b][p] += A[b][i] . XhT[p][i]             (M = 128, N = np, K = 64 rows, A from TMEM)
// Roles (ncu + in-kernel timers, profiles/README.md): a single issuing thread needs ~1000 clk per chunk for 5 MMAs,
// 3 waits and 2 commits (each wait ~120 clk even when satisfied, each MMA ~80 clk of dependent instructions) against
// 544 clk of tensor work, so MMA 1 and MMA 2 have an issuer thread each; a chunk's epilogue is a ~700 clk latency chain,
// so two groups of 8 warps take the chunks in turn (group = score / A buffer).  At the end the 128 x np accumulator
// is added to Cf with red.global (row ranges of different CTAs meet there).
// Measured (tools/umma_probe): the tensor core's fp32 accumulation truncates, ~1e-7 relative per instruction and
// always downwards, so a CTA accumulates at most VM_MAX_ROWS rows (1024 instructions) per accumulator.
#pragma once
#include "umma.cuh"

constexpr int VM_MAX_STAGES = 6, VM_STAGE_ROWS = 64, VM_CHUNK = 64, VM_THREADS = 640, VM_MAX_ROWS = 16384, VM_MAX_K16 = 4;
constexpr uint32_t VM_XT_BYTES = 256 * 128, VM_C8_BYTES = 128 * 64, VM_XL_BYTES = 64 * 32, VM_W_BYTES = 128 * 32;
// tensor-memory columns: E [0,256), scores D1 2 x 64 at 256, fp16 A operand 2 x 32 at 384
constexpr uint32_t VM_COL_E = 0, VM_COL_D1 = 256, VM_COL_A = 384;

struct VoteMmaParams {
  const double* wf;      // [nb][Ppad] final weights of the batch
  const double* inv_sd;  // [Ppad] 1 / global sd (the scaling of xh)
  const int* lv_off;
  const int* lv_k;
  const int* lv_blk;     // [L] first K = 16 block of the LV in the Xh_blk image
  float* Cf;             // [Ppad][L][ldl] (zeroed by the caller)
  const uint8_t* xt_img; // [chunk of 64 rows][p chunk] 32 KB
  const uint8_t* xl_img; // [chunk][K = 16 block] 2 KB
  const uint8_t* c8_img; // [chunk][tile of 128 replicates] 8 KB
  int n_blocks, n_rep_tiles_img;
  int64_t nb, ldl, N;
  int L, Ppad;
  int n_rep_tiles, n_pchunks, ksplit;
  int rows_per_cta;      // multiple of VM_STAGE_ROWS, <= VM_MAX_ROWS
  int k16_max;           // K = 16 steps of the widest block (<= VM_MAX_K16): sizes the ring stages
  int stages;            // ring depth (vm_stages(k16_max): as many as fit the shared memory, <= VM_MAX_STAGES)
  unsigned long long* stats;  // optional [16]: cycles waited per barrier, summed over CTAs (PLSPM_KERNEL_STATS)
};

__host__ __device__ inline uint32_t vm_stage_bytes(int k16_max) { return VM_XT_BYTES + VM_C8_BYTES + (uint32_t)k16_max * VM_XL_BYTES; }
// Ring depth: a stage is held from its bulk copies until the vote MMA of its chunk has completed (three chunk periods:
// score MMA, epilogue, vote MMA), so every stage beyond four is a copy in flight; take what the 227 KB allow.
__host__ inline int vm_stages(int k16_max, int max_smem = 227 * 1024) {
  static const int forced = getenv("PLSPM_VOTE_STAGES") ? atoi(getenv("PLSPM_VOTE_STAGES")) : 0;
  const int fit = (int)(((size_t)max_smem - 2048 - (size_t)k16_max * VM_W_BYTES) / vm_stage_bytes(k16_max));
  const int st = forced > 0 ? forced : fit;
  return st < 3 ? 3 : (st > VM_MAX_STAGES ? VM_MAX_STAGES : (st > fit ? fit : st));
}
__host__ inline size_t vm_smem_bytes(int k16_max, int stages) {
  return 1024 + (size_t)stages * vm_stage_bytes(k16_max) + (size_t)k16_max * VM_W_BYTES;
}

#define VM_TIMED(acc, ...)                                \
  {                                                       \
    const long long c_ = P.stats ? clock64() : 0;         \
    bar_wait_a(__VA_ARGS__);                              \
    if (P.stats) acc += clock64() - c_;                   \
  }

__global__ void __launch_bounds__(VM_THREADS, 1) vote_mma_kernel(const VoteMmaParams P) {
  using namespace umma;
  extern __shared__ uint8_t vm_smem_raw[];
  // barriers (8 bytes each): full[6] | empty[6] | d1_full[2] | d1_empty[2] | a_full[2] | e_full
  __shared__ uint64_t bars[19];
  __shared__ uint32_t tmem_base_sm;
  constexpr int B_FULL = 0, B_EMPTY = 6, B_D1F = 12, B_D1E = 14, B_AF = 16, B_EF = 18;
  const int S = P.stages;
  uint8_t* smem = vm_smem_raw + ((1024u - (s32(vm_smem_raw) & 1023u)) & 1023u);
  const uint32_t stage_bytes = vm_stage_bytes(P.k16_max);
  uint8_t* wtile = smem + (size_t)S * stage_bytes;  // [k16][128 x 16] fp16, no swizzle (8x16B core matrices)

  // tile of this CTA
  int t = blockIdx.x;
  const int rt = t % P.n_rep_tiles; t /= P.n_rep_tiles;
  const int l = t % P.L; t /= P.L;
  const int pc = t % P.n_pchunks;
  const int ks = t / P.n_pchunks;
  const int64_t b0 = (int64_t)rt * 128;
  const int p0 = pc * 256;
  const int np = min(256, (P.Ppad - p0 + 15) & ~15);  // MMA 2 N
  const int lvo = P.lv_off[l], lvk = P.lv_k[l];
  const int k16 = (lvk + 15) >> 4;
  const int64_t row_begin = (int64_t)ks * P.rows_per_cta;  // (a multiple of VM_STAGE_ROWS)
  const int64_t row_end = min(P.N, row_begin + P.rows_per_cta);
  const int n_chunks = row_end > row_begin ? (int)((row_end - row_begin + VM_CHUNK - 1) / VM_CHUNK) : 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 2) tmem_alloc(&tmem_base_sm, 512);
  if (threadIdx.x == 0) {
    for (int s = 0; s < VM_MAX_STAGES; ++s) { bar_init(&bars[B_FULL + s], 1); bar_init(&bars[B_EMPTY + s], 1); }
    for (int s = 0; s < 2; ++s) { bar_init(&bars[B_D1F + s], 1); bar_init(&bars[B_D1E + s], 8); bar_init(&bars[B_AF + s], 8); }
    bar_init(&bars[B_EF], 1);
    bar_fence_init();
  }
  // W_l: w'[b][q] = wf[b][lv_off + q] * sd_q for the block's columns, zero elsewhere (the K = 16 step may run into
  // the next block's columns or past the matrix: those products must vanish)
  for (int e = threadIdx.x; e < k16 * 128 * 16; e += VM_THREADS) {
    const int s = e >> 11, r = (e >> 4) & 127, c = e & 15, q = s * 16 + c;
    float v = 0.f;
    if (q < lvk && b0 + r < P.nb) {
      const double isd = P.inv_sd[lvo + q];
      v = isd > 0.0 ? (float)(P.wf[(b0 + r) * P.Ppad + lvo + q] / isd) : 0.f;
    }
    *reinterpret_cast<__half*>(wtile + (size_t)s * VM_W_BYTES + (r >> 3) * 256 + (c >> 3) * 128 + (r & 7) * 16 + (c & 7) * 2) = __float2half_rn(v);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tmem_base_sm;
  const uint32_t bar0 = s32(bars), smem0 = s32(smem);

  if (n_chunks > 0) {
    if (warp == 0) {
      if (lane == 0) {  // ---- producer: three linear bulk copies per chunk ------------------------------------------------
        const int64_t S0 = row_begin / VM_CHUNK;
        const uint8_t* xt = P.xt_img + ((size_t)S0 * P.n_pchunks + pc) * VM_XT_BYTES;
        const uint8_t* c8 = P.c8_img + ((size_t)S0 * P.n_rep_tiles_img + rt) * VM_C8_BYTES;
        const uint8_t* xl = P.xl_img + ((size_t)S0 * P.n_blocks + P.lv_blk[l]) * VM_XL_BYTES;
        const size_t xt_step = (size_t)P.n_pchunks * VM_XT_BYTES, c8_step = (size_t)P.n_rep_tiles_img * VM_C8_BYTES,
                     xl_step = (size_t)P.n_blocks * VM_XL_BYTES;
        const uint32_t xl_bytes = (uint32_t)k16 * VM_XL_BYTES, tx = VM_XT_BYTES + VM_C8_BYTES + xl_bytes;
        long long w1 = 0;
        for (int j = 0, s = 0, ph = 0; j < n_chunks; ++j, xt += xt_step, c8 += c8_step, xl += xl_step) {
          const uint32_t st = smem0 + (uint32_t)s * stage_bytes, fb = bar0 + 8 * (B_FULL + s);
          VM_TIMED(w1, bar0 + 8 * (B_EMPTY + s), ph ^ 1, 1);
          bar_expect_tx_a(fb, tx);
          bulk_load_a(st, xt, VM_XT_BYTES, fb);
          bulk_load_a(st + VM_XT_BYTES, c8, VM_C8_BYTES, fb);
          bulk_load_a(st + VM_XT_BYTES + VM_C8_BYTES, xl, xl_bytes, fb);
          if (++s == S) { s = 0; ph ^= 1; }
        }
        if (P.stats) atomicAdd(P.stats + 1, (unsigned long long)w1);
      }
    } else if (warp == 1) {
      if (lane == 0) {  // ---- score-MMA issuer ----------------------------------------------------------------------
        const uint32_t idesc1 = instr_desc(D_F32, AB_F16, AB_F16, 0, 0, 128, VM_CHUNK);
        const uint64_t wdesc = smem_desc(s32(wtile), 128, 256, SW_NONE);
        const uint64_t xdesc0 = smem_desc(smem0 + VM_XT_BYTES + VM_C8_BYTES, 16, 256, SW_32B);
        const uint32_t sdesc = stage_bytes >> 4;  // a stage further in descriptor units (start-address field only)
        const uint32_t d1_col = tbase + VM_COL_D1;
        long long w3 = 0, w4 = 0;
        const long long t_begin = clock64();
        for (int j = 0, s = 0, ph = 0; j < n_chunks; ++j) {
          const int u = j & 1;
          VM_TIMED(w3, bar0 + 8 * (B_FULL + s), ph, 3);
          VM_TIMED(w4, bar0 + 8 * (B_D1E + u), ((j >> 1) & 1) ^ 1, 4);
          tc_fence_after();
          const uint64_t xdesc = xdesc0 + (uint64_t)(sdesc * s);
          for (int q = 0; q < k16; ++q)
            mma_f16_ss(d1_col + 64 * u, wdesc + (uint64_t)(q * (VM_W_BYTES >> 4)), xdesc + (uint64_t)(q * (VM_XL_BYTES >> 4)), idesc1, q ? 1u : 0u);
          mma_commit_a(bar0 + 8 * (B_D1F + u));
          if (++s == S) { s = 0; ph ^= 1; }
        }
        if (P.stats) {
          atomicAdd(P.stats + 3, (unsigned long long)w3);
          atomicAdd(P.stats + 4, (unsigned long long)w4);
          atomicAdd(P.stats + 0, (unsigned long long)(clock64() - t_begin));
          atomicAdd(P.stats + 9, (unsigned long long)n_chunks);
        }
      }
    } else if (warp == 3) {
      if (lane == 0) {  // ---- vote-MMA issuer -------------------------------------------------------------------------
        const uint32_t idesc2 = instr_desc(D_F32, AB_F16, AB_F16, 0, 0, 128, (uint32_t)np);
        const uint64_t bdesc0 = smem_desc(smem0, 16, 1024, SW_128B);
        const uint32_t sdesc = stage_bytes >> 4;
        const uint32_t a_col = tbase + VM_COL_A, e_col = tbase + VM_COL_E;
        long long w2 = 0;
        const long long t_begin = clock64();
        for (int j = 0, s = 0; j < n_chunks; ++j) {
          const int u = j & 1;
          VM_TIMED(w2, bar0 + 8 * (B_AF + u), (j >> 1) & 1, 2);
          // (a_full of chunk j implies full[s]: the epilogue waited for it)
          tc_fence_after();
          const uint64_t bdesc = bdesc0 + (uint64_t)(sdesc * s);
#pragma unroll
          for (int k = 0; k < VM_CHUNK / 16; ++k)
            mma_f16_ts(e_col, a_col + 32 * u + 8 * k, bdesc + 2 * k, idesc2, (j | k) ? 1u : 0u);
          // the ring stage is free, and so is the A buffer (the epilogue of chunk j + 2 waits for the same event)
          mma_commit_a(bar0 + 8 * (B_EMPTY + s));
          if (++s == S) s = 0;
        }
        mma_commit_a(bar0 + 8 * B_EF);
        if (P.stats) {
          atomicAdd(P.stats + 2, (unsigned long long)w2);
          atomicAdd(P.stats + 8, (unsigned long long)(clock64() - t_begin));
        }
      }
    } else if (warp >= 4) {  // ---- epilogue warps: scores -> multiplicity-weighted fp16 A operand -----------------
      // two groups of 8 warps take the chunks in turn (group = chunk parity = score / A buffer index)
      const int grp = (warp - 4) >> 3;
      const int q = warp & 3, h = ((warp - 4) >> 2) & 1;
      const int r = 32 * q + lane;  // replicate row of this thread = TMEM lane
      const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16);
      // this thread's two 16-byte chunks of its multiplicity row (64 B) inside a stage (64B-swizzled [128 x 64] tile)
      const int sw = (r >> 1) & 3;
      const uint32_t c8a = (uint32_t)VM_XT_BYTES + (uint32_t)r * 64 + (uint32_t)(((2 * h) ^ sw) << 4);
      const uint32_t c8b = (uint32_t)VM_XT_BYTES + (uint32_t)r * 64 + (uint32_t)(((2 * h + 1) ^ sw) << 4);
      const int u = grp;
      long long w5 = 0, w7 = 0;
      for (int j = grp; j < n_chunks; j += 2) {
        const int s = j % S, n = j >> 1;  // n: use count of this group's buffers
        VM_TIMED(w5, bar0 + 8 * (B_D1F + u), n & 1, 5);
        tc_fence_after();
        uint32_t tt[32];
        tmem_ld32(lane_addr + VM_COL_D1 + 64 * u + 32 * h, tt);
        bar_wait_a(bar0 + 8 * (B_FULL + s), (j / S) & 1, 6);  // (complete long ago: acquires the bulk-copy writes)
        const uint8_t* st = smem + (size_t)s * stage_bytes;
        const uint4 ca = *reinterpret_cast<const uint4*>(st + c8a);
        const uint4 cb = *reinterpret_cast<const uint4*>(st + c8b);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive_a(bar0 + 8 * (B_D1E + u));
        const uint32_t cw[8] = {ca.x, ca.y, ca.z, ca.w, cb.x, cb.y, cb.z, cb.w};
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          // byte -> float without a conversion instruction: 0x4B0000bb is 2^23 + bb
          const float c0 = __uint_as_float(__byte_perm(cw[e >> 2], 0x4B000000u, 0x7440 | (e & 3))) - 8388608.f;
          const float c1 = __uint_as_float(__byte_perm(cw[e >> 2], 0x4B000000u, 0x7440 | ((e + 1) & 3))) - 8388608.f;
          const __half2 h2 = __floats2half2_rn(c0 * __uint_as_float(tt[e]), c1 * __uint_as_float(tt[e + 1]));
          pk[e >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
        }
        // MMA 2 of this group's previous chunk (j - 2) has read the A buffer: the event that released its ring stage
        if (j >= 2) VM_TIMED(w7, bar0 + 8 * (B_EMPTY + ((j - 2) % S)), ((j - 2) / S) & 1, 7);
        tc_fence_after();
        tmem_st16(lane_addr + VM_COL_A + 32 * u + 16 * h, pk);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive_a(bar0 + 8 * (B_AF + u));
      }
      if (P.stats && (threadIdx.x == 128)) {
        atomicAdd(P.stats + 5, (unsigned long long)w5);
        atomicAdd(P.stats + 7, (unsigned long long)w7);
      }
      // ---- the accumulator of this row range joins the others in Cf --------------------------------------------
      bar_wait_a(bar0 + 8 * B_EF, 0, 8);
      tc_fence_after();
      const int64_t b = b0 + r;
      for (int c0 = 128 * h + 64 * grp; c0 < min(np, 128 * h + 64 * grp + 64); c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + VM_COL_E + c0, v);
        tmem_wait_ld();
        if (b < P.nb) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int p = p0 + c0 + c;
            if (p < P.Ppad) atomicAdd(P.Cf + ((size_t)p * P.L + l) * P.ldl + b, __uint_as_float(v[c]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tbase, 512);
}

// ---- operand images (built once per data handle) -------------------------------------------------------------------
// XhT image: xh = fp16(x~ / sd), K-major.  Image (chunk S of 64 rows, column chunk pc of 256) = 32 KB at
// ((S * n_pchunks + pc) * 32 KB) = a [256 p x 64 k] tile with the 128-byte swizzle: element (p, k) at
// p * 128 + (((k >> 3) ^ (p & 7)) << 4) + (k & 7) * 2.  Columns >= Ppad and rows >= N are zero.
// One thread = one 16-byte chunk (8 consecutive rows of one column).
__global__ void __launch_bounds__(256) vote_xt_image_kernel(const double* __restrict__ X, int64_t N, int Ppad, int n_pchunks,
                                                            int64_t n_chunks, const double* __restrict__ inv_sd,
                                                            uint8_t* __restrict__ img) {
  const int64_t total = n_chunks * n_pchunks * 256 * 8;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    // consecutive threads: consecutive columns p (coalesced reads of the row-major x~), then the 8 chunks, then images
    const int pl = (int)(e & 255), ch = (int)((e >> 8) & 7);
    const int64_t im = e >> 11, S = im / n_pchunks;
    const int p = (int)(im - S * n_pchunks) * 256 + pl;
    const int64_t i0 = S * 64 + 8 * ch;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (p < Ppad) {
      const double isd = inv_sd[p];
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const float a = (i0 + k < N) ? (float)(X[(i0 + k) * Ppad + p] * isd) : 0.f;
        const float b = (i0 + k + 1 < N) ? (float)(X[(i0 + k + 1) * Ppad + p] * isd) : 0.f;
        const __half2 h2 = __floats2half2_rn(a, b);
        w[k >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
      }
    }
    *reinterpret_cast<uint4*>(img + (size_t)im * 32768 + (size_t)pl * 128 + ((ch ^ (pl & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// Xh_blk image: the block columns of xh for the score MMA.  Image (chunk S, K = 16 block u) = 2 KB at
// ((S * n_blocks + u) * 2 KB) = [64 rows x 16 columns] fp16 with the 32-byte swizzle: element (row r, column j) at
// r * 32 + (((j >> 3) ^ ((r >> 2) & 1)) << 4) + (j & 7) * 2; block u covers padded columns blk_col[u] .. + 15.
__global__ void __launch_bounds__(256) vote_xl_image_kernel(const double* __restrict__ X, int64_t N, int Ppad, int n_blocks,
                                                            int64_t n_chunks, const int* __restrict__ blk_col,
                                                            const double* __restrict__ inv_sd, uint8_t* __restrict__ img) {
  const int64_t total = n_chunks * n_blocks * 64 * 2;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int hb = (int)(e & 1), r = (int)((e >> 1) & 63);
    const int64_t im = e >> 7, S = im / n_blocks;
    const int u = (int)(im - S * n_blocks);
    const int64_t i = S * 64 + r;
    const int c0 = blk_col[u] + 8 * hb;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (i < N) {
#pragma unroll
      for (int k = 0; k < 8; k += 2) {
        const float a = (c0 + k < Ppad) ? (float)(X[i * Ppad + c0 + k] * inv_sd[c0 + k]) : 0.f;
        const float b = (c0 + k + 1 < Ppad) ? (float)(X[i * Ppad + c0 + k + 1] * inv_sd[c0 + k + 1]) : 0.f;
        const __half2 h2 = __floats2half2_rn(a, b);
        w[k >> 1] = *reinterpret_cast<const uint32_t*>(&h2);
      }
    }
    *reinterpret_cast<uint4*>(img + (size_t)im * 2048 + (size_t)r * 32 + ((hb ^ ((r >> 2) & 1)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
// The multiplicities for the vote kernel: image (chunk S of 64 rows, tile rt of 128 replicates) = 8 KB at
// ((S * n_rep_tiles + rt) * 8 KB) = [128 replicates x 64 B] with the 64-byte swizzle: byte (replicate r, row k) at
// r * 64 + (((k >> 4) ^ ((r >> 1) & 3)) << 4) + (k & 15).  (The Gram kernel has its own image layout of the same numbers.)
__global__ void __launch_bounds__(256) vote_c8_image_kernel(const uint32_t* __restrict__ counts, int64_t N, int64_t nb, int n_rep_tiles,
                                                            int64_t n_chunks, uint8_t* __restrict__ img) {
  const int64_t total = n_chunks * n_rep_tiles * 128 * 4;  // 16-byte chunks
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(e & 3), r = (int)((e >> 2) & 127);
    const int64_t im = e >> 9, S = im / n_rep_tiles;
    const int64_t b = (im - S * n_rep_tiles) * 128 + r, i0 = S * 64 + 16 * c;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
    if (b < nb && i0 < N) {
      const uint32_t* src = counts + b * N + i0;
#pragma unroll
      for (int k = 0; k < 16; ++k) w[k >> 2] |= ((i0 + k < N ? src[k] : 0u) & 127u) << (8 * (k & 3));
    }
    *reinterpret_cast<uint4*>(img + (size_t)im * 8192 + (size_t)r * 64 + ((c ^ ((r >> 1) & 3)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}
