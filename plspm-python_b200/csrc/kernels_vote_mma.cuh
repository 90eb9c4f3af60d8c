// Fused tensor-core sign vote (tcgen05 + TMEM + TMA): replaces score generation + a library GEMM.
// Part of the single translation unit plspm_b200.cu (included there); see DESIGN.md §4.
//
//   E'[p][l][b] = sum_i xh_ip * fp16( c_bi * t'_bil ),     t'_bil = sum_{q in block l} xh_iq w'_bq,  w' = wf * sd
//
// (xh = x~/sd in fp16; t' is the UN-centred score: the solver removes sh_l * sum_i c_bi xh_ip afterwards, see
// solver_core.h phase 3.)  Only the SIGN of E is used, and only where it exceeds a rigorous error bound.
//
// One CTA = (128 replicates, one latent variable l, <= 256 manifest columns p, one range of rows).  Per chunk of
// 64 rows, everything stays on the SM:
//   TMA       XhT tile [np x 64] (K-major, 128B swizzle), the block columns of Xh [64 x 16] (32B swizzle) and the
//             multiplicities c8 [128 x 64] (64B swizzle) land in a 4-stage shared-memory ring
//   MMA 1     D1[b][i] = W_l[b][:] . Xh_blk[i][:]       (M = 128 replicates, N = 64 rows, K = 16 per step, SS)
//             -> the scores of 128 replicates x 64 rows, fp32, in tensor memory
//   epilogue  8 warps: tcgen05.ld the scores, multiply by the row multiplicity, round to fp16, tcgen05.st them
//             back to tensor memory as the A operand of
//   MMA 2     E[b][p] += A[b][i] . XhT[p][i]             (M = 128, N = np, K = 64 rows, A from TMEM)
// MMA 1 of chunk j+1 is issued before MMA 2 of chunk j, so the epilogue of one chunk overlaps the vote MMAs of the
// previous one; the scores never touch shared or global memory.  At the end the 128 x np accumulator is added to
// Cf with red.global (row ranges of different CTAs meet there).
// Measured (tools/umma_probe): the tensor core's fp32 accumulation truncates, ~1e-7 relative per instruction and
// always downwards, so a CTA accumulates at most VM_MAX_ROWS rows (1024 instructions) per accumulator.
#pragma once
#include "umma.cuh"

constexpr int VM_STAGES = 4, VM_CHUNK = 64, VM_THREADS = 384, VM_MAX_ROWS = 16384, VM_MAX_K16 = 4;
constexpr uint32_t VM_XT_BYTES = 256 * 128, VM_C8_BYTES = 128 * 64, VM_XL_BYTES = 64 * 32, VM_W_BYTES = 128 * 32;
// tensor-memory columns: E [0,256), scores D1 2 x 64 at 256, fp16 A operand 2 x 32 at 384
constexpr uint32_t VM_COL_E = 0, VM_COL_D1 = 256, VM_COL_A = 384;

struct VoteMmaParams {
  const double* wf;      // [nb][Ppad] final weights of the batch
  const double* inv_sd;  // [Ppad] 1 / global sd (the scaling of xh)
  const int* lv_off;
  const int* lv_k;
  float* Cf;             // [Ppad][L][ldl] (zeroed by the caller)
  int64_t nb, ldl, N;
  int L, Ppad;
  int n_rep_tiles, n_pchunks, ksplit;
  int rows_per_cta;      // multiple of VM_CHUNK, <= VM_MAX_ROWS
  int k16_max;           // K = 16 steps of the widest block (<= VM_MAX_K16): sizes the ring stages
  int np_box;            // rows of the XhT TMA box (min(256, Ppad rounded up to 16)): a box is always delivered whole
};

__host__ __device__ inline uint32_t vm_stage_bytes(int k16_max) { return VM_XT_BYTES + VM_C8_BYTES + (uint32_t)k16_max * VM_XL_BYTES; }
__host__ inline size_t vm_smem_bytes(int k16_max) {
  return 1024 + (size_t)VM_STAGES * vm_stage_bytes(k16_max) + (size_t)k16_max * VM_W_BYTES;
}

__global__ void __launch_bounds__(VM_THREADS, 1)
    vote_mma_kernel(const __grid_constant__ CUtensorMap map_xt, const __grid_constant__ CUtensorMap map_xh,
                    const __grid_constant__ CUtensorMap map_c8, const VoteMmaParams P) {
  using namespace umma;
  extern __shared__ uint8_t vm_smem_raw[];
  __shared__ uint64_t full[VM_STAGES], empty[VM_STAGES], d1_full[2], d1_empty[2], a_full[2], a_empty[2], e_full;
  __shared__ uint32_t tmem_base_sm;
  uint8_t* smem = vm_smem_raw + ((1024u - (s32(vm_smem_raw) & 1023u)) & 1023u);
  const uint32_t stage_bytes = vm_stage_bytes(P.k16_max);
  uint8_t* wtile = smem + (size_t)VM_STAGES * stage_bytes;  // [k16][128 x 16] fp16, no swizzle (8x16B core matrices)

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
  const int64_t row_begin = (int64_t)ks * P.rows_per_cta;
  const int64_t row_end = min(P.N, row_begin + P.rows_per_cta);
  const int n_chunks = row_end > row_begin ? (int)((row_end - row_begin + VM_CHUNK - 1) / VM_CHUNK) : 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 2) tmem_alloc(&tmem_base_sm, 512);
  if (threadIdx.x == 0) {
    for (int s = 0; s < VM_STAGES; ++s) { bar_init(&full[s], 1); bar_init(&empty[s], 1); }
    for (int u = 0; u < 2; ++u) { bar_init(&d1_full[u], 1); bar_init(&d1_empty[u], 8); bar_init(&a_full[u], 8); bar_init(&a_empty[u], 1); }
    bar_init(&e_full, 1);
    bar_fence_init();
    tma_prefetch_desc(&map_xt);
    tma_prefetch_desc(&map_xh);
    tma_prefetch_desc(&map_c8);
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

  if (n_chunks > 0) {
    if (warp == 0) {
      if (lane == 0) {  // ---- TMA producer --------------------------------------------------------------------
        // (out-of-bounds parts of a box arrive as zeros and count: the byte total never depends on the position)
        const uint32_t tx = (uint32_t)P.np_box * 128u + VM_C8_BYTES + (uint32_t)k16 * VM_XL_BYTES;
        for (int j = 0; j < n_chunks; ++j) {
          const int s = j % VM_STAGES;
          bar_wait(&empty[s], ((j / VM_STAGES) & 1) ^ 1, 1);
          uint8_t* st = smem + (size_t)s * stage_bytes;
          const int i0 = (int)(row_begin + (int64_t)j * VM_CHUNK);
          bar_expect_tx(&full[s], tx);
          tma_load_2d(st, &map_xt, &full[s], i0, p0);
          tma_load_2d(st + VM_XT_BYTES, &map_c8, &full[s], i0, (int)b0);
          for (int q = 0; q < k16; ++q) tma_load_2d(st + VM_XT_BYTES + VM_C8_BYTES + q * VM_XL_BYTES, &map_xh, &full[s], lvo + 16 * q, i0);
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {  // ---- MMA issuer ----------------------------------------------------------------------
        const uint32_t idesc1 = instr_desc(D_F32, AB_F16, AB_F16, 0, 0, 128, VM_CHUNK);
        const uint32_t idesc2 = instr_desc(D_F32, AB_F16, AB_F16, 0, 0, 128, (uint32_t)np);
        const uint64_t wdesc = smem_desc(s32(wtile), 128, 256, SW_NONE);
        auto vote_mma = [&](int j) {  // MMA 2 of chunk j
          const int s = j % VM_STAGES, u = j & 1;
          bar_wait(&a_full[u], (j >> 1) & 1, 2);
          tc_fence_after();
          const uint64_t bdesc = smem_desc(s32(smem + (size_t)s * stage_bytes), 16, 1024, SW_128B);
#pragma unroll
          for (int k = 0; k < VM_CHUNK / 16; ++k)
            mma_f16_ts(tbase + VM_COL_E, tbase + VM_COL_A + 32 * u + 8 * k, desc_advance(bdesc, 32 * k), idesc2, (j | k) ? 1u : 0u);
          mma_commit(&empty[s]);
          mma_commit(&a_empty[u]);
        };
        for (int j = 0; j < n_chunks; ++j) {
          const int s = j % VM_STAGES, u = j & 1;
          bar_wait(&full[s], (j / VM_STAGES) & 1, 3);
          bar_wait(&d1_empty[u], ((j >> 1) & 1) ^ 1, 4);
          tc_fence_after();
          const uint64_t xdesc = smem_desc(s32(smem + (size_t)s * stage_bytes + VM_XT_BYTES + VM_C8_BYTES), 16, 256, SW_32B);
          for (int q = 0; q < k16; ++q)
            mma_f16_ss(tbase + VM_COL_D1 + 64 * u, desc_advance(wdesc, q * VM_W_BYTES), desc_advance(xdesc, q * VM_XL_BYTES), idesc1, q ? 1u : 0u);
          mma_commit(&d1_full[u]);
          if (j > 0) vote_mma(j - 1);
        }
        vote_mma(n_chunks - 1);
        mma_commit(&e_full);
      }
    } else if (warp >= 4) {  // ---- epilogue warps: scores -> multiplicity-weighted fp16 A operand -----------------
      const int q = warp & 3, h = (warp - 4) >> 2;
      const int r = 32 * q + lane;  // replicate row of this thread = TMEM lane
      const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16);
      for (int j = 0; j < n_chunks; ++j) {
        const int s = j % VM_STAGES, u = j & 1;
        bar_wait(&d1_full[u], (j >> 1) & 1, 5);
        tc_fence_after();
        uint32_t tt[32];
        tmem_ld32(lane_addr + VM_COL_D1 + 64 * u + 32 * h, tt);
        bar_wait(&full[s], (j / VM_STAGES) & 1, 6);  // (complete long ago: acquires the TMA writes for this thread)
        const uint8_t* c8 = smem + (size_t)s * stage_bytes + VM_XT_BYTES + (size_t)r * 64;
        const int sw = (r >> 1) & 3;
        const uint4 ca = *reinterpret_cast<const uint4*>(c8 + (((2 * h) ^ sw) << 4));
        const uint4 cb = *reinterpret_cast<const uint4*>(c8 + (((2 * h + 1) ^ sw) << 4));
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&d1_empty[u]);
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
        bar_wait(&a_empty[u], ((j >> 1) & 1) ^ 1, 7);
        tc_fence_after();
        tmem_st16(lane_addr + VM_COL_A + 32 * u + 16 * h, pk);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(&a_full[u]);
      }
      // ---- the accumulator of this row range joins the others in Cf --------------------------------------------
      bar_wait(&e_full, 0, 8);
      tc_fence_after();
      const int64_t b = b0 + r;
      for (int c0 = 128 * h; c0 < min(np, 128 * h + 128); c0 += 32) {
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

// xhT[p][i] = fp16(x~_ip * inv_sd_p): the K-major B operand of the vote MMA (one 32 x 32 tile per block, via shared memory)
__global__ void __launch_bounds__(256) make_half_t_kernel(const double* __restrict__ X, int64_t N, int Ppad, int64_t ldt,
                                                          const double* __restrict__ inv_sd, __half* __restrict__ XhT) {
  __shared__ __half tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int p0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int rr = ty; rr < 32; rr += 8) {
    const int64_t i = i0 + rr;
    const int p = p0 + tx;
    tile[rr][tx] = (i < N && p < Ppad) ? __double2half(X[i * Ppad + p] * inv_sd[p]) : __float2half(0.f);
  }
  __syncthreads();
  for (int rr = ty; rr < 32; rr += 8) {
    const int p = p0 + rr;
    const int64_t i = i0 + tx;
    if (p < Ppad && i < ldt) XhT[(int64_t)p * ldt + i] = tile[tx][rr];
  }
}
