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

__global__ void indices_kernel(int32_t* __restrict__ out, int64_t N, uint64_t rep, uint64_t seed) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= (N + 3) / 4) return;
  uint32_t r[4];
  philox4x32_10((uint32_t)g, (uint32_t)((uint64_t)g >> 32), (uint32_t)rep, (uint32_t)(rep >> 32), (uint32_t)seed,
                (uint32_t)(seed >> 32), r);
  for (int k = 0; k < 4; ++k)
    if (g * 4 + k < N) out[g * 4 + k] = (int32_t)index_from_u32(r[k], (uint32_t)N);
}
