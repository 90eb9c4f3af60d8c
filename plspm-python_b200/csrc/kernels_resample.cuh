// Philox4x32-10 resampling: multiplicity histograms and index lists.
// Part of the single translation unit plspm_b200.cu (included there, in this order); see DESIGN.md §4.
#pragma once

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Random123); counter = (row group, 0, replicate lo, replicate hi), key = seed
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                      uint32_t k1, uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
__host__ __device__ __forceinline__ uint32_t index_from_u32(uint32_t u, uint32_t N) {
  return (uint32_t)(((uint64_t)u * (uint64_t)N) >> 32);
}

// counts[b][i] += multiplicity of row i in replicate b.  One thread = 4 consecutive draws.
__global__ void counts_kernel(uint32_t* __restrict__ counts, const int32_t* __restrict__ idx, int64_t N, int64_t nrep,
                              int64_t rep_begin, uint64_t seed) {
  const int64_t groups = (N + 3) / 4;
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= groups * nrep) return;
  const int64_t b = gid / groups, g = gid - b * groups;
  uint32_t* c = counts + b * N;
  if (idx) {
    const int32_t* ib = idx + b * N;
    for (int k = 0; k < 4; ++k) {
      int64_t i = g * 4 + k;
      if (i < N) atomicAdd(&c[(uint32_t)ib[i]], 1u);
    }
  } else {
    uint64_t rep = (uint64_t)(rep_begin + b);
    uint32_t r[4];
    philox4x32_10((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed,
                  (uint32_t)(seed >> 32), r);
    for (int k = 0; k < 4; ++k)
      if (g * 4 + k < N) atomicAdd(&c[index_from_u32(r[k], (uint32_t)N)], 1u);
  }
}

// Bootstrap multiplicities of a batch written DIRECTLY as the int8 shared-memory tile images the two tcgen05 kernels
// fetch (layouts: counts8_image_kernel in kernels_gram_mma.cuh, vote_c8_image_kernel in kernels_vote_mma.cuh), without
// the [nb x N] uint32 table in HBM (614 MB written, read twice per c3 batch).  CTA = (replicate b, row range): the
// draws of the replicate -- the same Philox groups as counts_kernel, so the same resamples -- are counted in shared
// memory as packed bytes (atomicAdd of 1 << 8*(i & 3) on the word), then stored 16 bytes at a time into both images.
// A byte reaching 128 is flagged BEFORE it can carry into its neighbour (the previous value is returned by the atomic;
// every increment past 127 sees it), and the flagged batch is redone on the fp64 route from a uint32 table.
// Replicates >= nb (image padding) are written as zeros.  Rows beyond N are zero.
#define RI_THREADS 512
#define RI_MAX_ROWS 204800
__global__ void __launch_bounds__(RI_THREADS) resample_images_kernel(const int32_t* __restrict__ idx, int64_t N, int64_t nb,
                                                                     int64_t rep_begin, uint64_t seed, int64_t range_rows,
                                                                     int n_groups, int n_rep_tiles, int64_t n_chunks64,
                                                                     uint8_t* __restrict__ gimg, uint8_t* __restrict__ vimg,
                                                                     int* __restrict__ overflow) {
  extern __shared__ uint32_t ri_words[];
  const int64_t b = blockIdx.x, r0 = (int64_t)blockIdx.y * range_rows;
  const int64_t n_pad = (N + 127) / 128 * 128;
  const int64_t rows = min(range_rows, n_pad - r0);
  for (int64_t e = threadIdx.x; e < rows / 4; e += RI_THREADS) ri_words[e] = 0u;
  __syncthreads();
  if (b < nb) {
    const int64_t groups = (N + 3) / 4;
    const uint64_t rep = (uint64_t)(rep_begin + b);
    bool big = false;
    for (int64_t g = threadIdx.x; g < groups; g += RI_THREADS) {
      uint32_t r[4];
      if (idx) {
        for (int k = 0; k < 4; ++k) r[k] = g * 4 + k < N ? (uint32_t)idx[b * N + g * 4 + k] : 0u;
      } else {
        philox4x32_10((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        for (int k = 0; k < 4; ++k) r[k] = index_from_u32(r[k], (uint32_t)N);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int64_t rel = (int64_t)r[k] - r0;
        if (g * 4 + k < N && (uint64_t)rel < (uint64_t)rows) {
          const int sh = 8 * (int)(rel & 3);
          const uint32_t old = atomicAdd(&ri_words[rel >> 2], 1u << sh);
          big |= ((old >> sh) & 0xffu) >= 127u;
        }
      }
    }
    if (big) *overflow = 1;
  }
  __syncthreads();
  const int64_t G = b >> 9;
  const int rg = (int)(b & 511), rv = (int)(b & 127);
  const int64_t T = b >> 7;
  const uint4* src = reinterpret_cast<const uint4*>(ri_words);
  for (int64_t e = threadIdx.x; e < rows / 16; e += RI_THREADS) {
    const int64_t c16 = r0 / 16 + e;  // 16-row chunk of the data
    const uint4 v = src[e];
    {
      const int64_t S = c16 >> 3;
      const int c = (int)(c16 & 7);
      *reinterpret_cast<uint4*>(gimg + (size_t)(S * n_groups + G) * 65536 + (size_t)(rg >> 8) * 32768 + (size_t)(rg & 255) * 128 +
                                ((c ^ (rg & 7)) << 4)) = v;
    }
    const int64_t S = c16 >> 2;
    if (T < n_rep_tiles && S < n_chunks64) {
      const int c = (int)(c16 & 3);
      *reinterpret_cast<uint4*>(vimg + (size_t)(S * n_rep_tiles + T) * 8192 + (size_t)rv * 64 + ((c ^ ((rv >> 1) & 3)) << 4)) = v;
    }
  }
}

__global__ void indices_kernel(int32_t* __restrict__ out, int64_t N, uint64_t rep, uint64_t seed) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (N + 3) / 4) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  for (int k = 0; k < 4; ++k)
    if (g * 4 + k < N) out[g * 4 + k] = (int32_t)index_from_u32(r[k], (uint32_t)N);
}
