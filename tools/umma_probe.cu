// Validates the tcgen05 / TMEM / TMA building blocks of plspm-python_b200/csrc/umma.cuh on a B200, one
// experiment per process invocation (a wrong descriptor can poison the context):
//   umma_probe tmem                         TMEM st/ld round trip, column / lane addressing
//   umma_probe dump <sw> <esize> <boxc> <boxr>   raw image of a TMA tile with swizzle sw in {0,32,64,128}
//   umma_probe gemm <name> [key=value ...]  one UMMA tile product against a CPU reference
// tools/run_probe.sh runs the list of experiments the kernels rely on.
#include <cuda_fp16.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <random>
#include <string>
#include <vector>

#include "../plspm-python_b200/csrc/umma.cuh"
using namespace umma;

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at %s:%d: %s\n", #x, __FILE__, __LINE__, cudaGetErrorString(e_)); \
      exit(2);                                                                                 \
    }                                                                                          \
  } while (0)

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tmem(uint32_t* out) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tbase, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tbase;
  uint32_t r[32], q[32], q8[8];
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x * 1000u + i;
  const uint32_t addr = tmem_addr(base, 32u * warp, 0);
  tmem_st32(addr, r);
  tmem_wait_st();
  tmem_ld32(addr, q);
  tmem_ld8(tmem_addr(base, 32u * warp, 8), q8);  // columns 8..15
  tmem_wait_ld();
  for (int i = 0; i < 32; ++i) out[threadIdx.x * 40 + i] = q[i];
  for (int i = 0; i < 8; ++i) out[threadIdx.x * 40 + 32 + i] = q8[i];
  if (threadIdx.x == 0) out[128 * 40] = base;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 64);
}

static int run_tmem() {
  uint32_t* d;
  CK(cudaMalloc(&d, (128 * 40 + 1) * 4));
  k_tmem<<<1, 128>>>(d);
  CK(cudaDeviceSynchronize());
  std::vector<uint32_t> h(128 * 40 + 1);
  CK(cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int t = 0; t < 128; ++t) {
    for (int i = 0; i < 32; ++i) bad += h[t * 40 + i] != t * 1000u + i;
    for (int i = 0; i < 8; ++i) bad += h[t * 40 + 32 + i] != t * 1000u + 8 + i;
  }
  printf("tmem base=0x%08x mismatches=%d %s\n", h[128 * 40], bad, bad ? "FAIL" : "PASS");
  return bad != 0;
}

// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_dump(const __grid_constant__ CUtensorMap map, uint32_t bytes, uint8_t* out) {
  extern __shared__ uint8_t dsm_raw[];
  __shared__ uint64_t bar;
  uint8_t* tile = dsm_raw + ((1024u - (s32(dsm_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0) {
    bar_init(&bar, 1);
    bar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    bar_expect_tx(&bar, bytes);
    tma_load_2d(tile, &map, &bar, 0, 0);
  }
  bar_wait(&bar, 0);
  for (uint32_t e = threadIdx.x; e < bytes; e += blockDim.x) out[e] = tile[e];
}

// position of byte j of row r in a swizzled tile with rows of W bytes (W = swizzle span: 32, 64 or 128)
static inline size_t swz(int sw, int W, int r, int j) {
  int chunk = j / 16;
  if (sw == 128) chunk ^= (r & 7);
  else if (sw == 64) chunk ^= ((r >> 1) & 3);
  else if (sw == 32) chunk ^= ((r >> 2) & 1);
  return (size_t)r * W + chunk * 16 + j % 16;
}

static int run_dump(int sw, int esize, int boxc, int boxr) {
  const int W = boxc * esize;  // bytes per box row
  const int cols = boxc * 2, rows = boxr + 3;
  std::vector<uint8_t> src((size_t)rows * cols * esize);
  for (size_t i = 0; i < src.size(); ++i) src[i] = (uint8_t)((i * 2654435761u) >> 13);
  uint8_t *dsrc, *dout;
  CK(cudaMalloc(&dsrc, src.size()));
  CK(cudaMemcpy(dsrc, src.data(), src.size(), cudaMemcpyHostToDevice));
  const uint32_t bytes = (uint32_t)W * boxr;
  CK(cudaMalloc(&dout, bytes));
  CUtensorMap map;
  CUtensorMapSwizzle s = sw == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : sw == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                         : sw == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUtensorMapDataType dt = esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16
                           : esize == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT64;
  if (!make_map_2d(&map, dsrc, dt, cols, rows, (uint64_t)cols * esize, boxc, boxr, s)) {
    printf("dump sw=%d: cuTensorMapEncodeTiled failed FAIL\n", sw);
    return 1;
  }
  CK(cudaFuncSetAttribute(k_dump, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes + 1024));
  k_dump<<<1, 128, bytes + 1024>>>(map, bytes, dout);
  CK(cudaDeviceSynchronize());
  std::vector<uint8_t> h(bytes);
  CK(cudaMemcpy(h.data(), dout, bytes, cudaMemcpyDeviceToHost));
  int bad = 0;
  for (int r = 0; r < boxr; ++r)
    for (int j = 0; j < W; ++j) bad += h[sw ? swz(sw, W, r, j) : (size_t)r * W + j] != src[(size_t)r * cols * esize + j];
  printf("dump sw=%d esize=%d box=%dx%d mismatches=%d %s\n", sw, esize, boxc, boxr, bad, bad ? "FAIL" : "PASS");
  return bad != 0;
}

// ---------------------------------------------------------------------------------------------------------
struct GemmArgs {
  int kind;  // 0: f16 (fp32 accumulate), 2: i8 (s32 accumulate)
  int N, ksteps, repeat;
  uint32_t a_bytes, b_bytes;
  uint32_t a_lbo, a_sbo, a_sw, a_kadv;
  uint32_t b_lbo, b_sbo, b_sw, b_kadv;
  uint32_t idesc;
  int a_src, b_src;  // 0: TMA tile (one box at the origin), 1: byte image prepared by the host, 2 (A only): TMEM
  int a_tmem_words;  // a_src == 2: 32-bit columns per row
};

__global__ void __launch_bounds__(128) k_gemm(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                              GemmArgs g, const uint8_t* __restrict__ a_img, const uint8_t* __restrict__ b_img,
                                              uint32_t* __restrict__ D) {
  extern __shared__ uint8_t gsm_raw[];
  __shared__ uint64_t bar_tma, bar_mma;
  __shared__ uint32_t tbase;
  uint8_t* sA = gsm_raw + ((1024u - (s32(gsm_raw) & 1023u)) & 1023u);
  uint8_t* sB = sA + ((g.a_bytes + 1023u) & ~1023u);
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&tbase, 512);
  if (threadIdx.x == 0) {
    bar_init(&bar_tma, 1);
    bar_init(&bar_mma, 1);
    bar_fence_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tbase;
  if (g.a_src == 1)
    for (uint32_t e = threadIdx.x; e < g.a_bytes / 16; e += blockDim.x) reinterpret_cast<uint4*>(sA)[e] = reinterpret_cast<const uint4*>(a_img)[e];
  if (g.b_src == 1)
    for (uint32_t e = threadIdx.x; e < g.b_bytes / 16; e += blockDim.x) reinterpret_cast<uint4*>(sB)[e] = reinterpret_cast<const uint4*>(b_img)[e];
  if (g.a_src == 2) {  // A in tensor memory at columns [256, 256 + a_tmem_words): lane = row
    const uint32_t* row = reinterpret_cast<const uint32_t*>(a_img) + (size_t)threadIdx.x * g.a_tmem_words;
    for (int c = 0; c < g.a_tmem_words; c += 8) {
      uint32_t r[8];
      for (int i = 0; i < 8; ++i) r[i] = row[c + i];
      tmem_st8(tmem_addr(base, 32u * warp, 256 + c), r);
    }
    tmem_wait_st();
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    uint32_t tx = 0;
    if (g.a_src == 0) tx += g.a_bytes;
    if (g.b_src == 0) tx += g.b_bytes;
    if (tx) {
      bar_expect_tx(&bar_tma, tx);
      if (g.a_src == 0) tma_load_2d(sA, &mapA, &bar_tma, 0, 0);
      if (g.b_src == 0) tma_load_2d(sB, &mapB, &bar_tma, 0, 0);
      bar_wait(&bar_tma, 0);
    }
    tc_fence_after();
    const uint64_t ad0 = smem_desc(s32(sA), g.a_lbo, g.a_sbo, g.a_sw), bd0 = smem_desc(s32(sB), g.b_lbo, g.b_sbo, g.b_sw);
    for (int rep = 0; rep < g.repeat; ++rep)
      for (int k = 0; k < g.ksteps; ++k) {
        const uint32_t acc = (rep | k) ? 1u : 0u;
        const uint64_t bd = desc_advance(bd0, g.b_kadv * k);
        if (g.a_src == 2) mma_f16_ts(base, tmem_addr(base, 0, 256 + 8 * k), bd, g.idesc, acc);
        else if (g.kind == 0) mma_f16_ss(base, desc_advance(ad0, g.a_kadv * k), bd, g.idesc, acc);
        else mma_i8_ss(base, desc_advance(ad0, g.a_kadv * k), bd, g.idesc, acc);
      }
    mma_commit(&bar_mma);
  }
  bar_wait(&bar_mma, 0);
  tc_fence_after();
  for (int c = 0; c < g.N; c += 8) {
    uint32_t r[8];
    tmem_ld8(tmem_addr(base, 32u * warp, c), r);
    tmem_wait_ld();
    for (int i = 0; i < 8; ++i) D[(size_t)threadIdx.x * g.N + c + i] = r[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

struct Opts {
  std::map<std::string, long> kv;
  long get(const char* k, long def) const {
    auto it = kv.find(k);
    return it == kv.end() ? def : it->second;
  }
};

// A: [128][K] (row m), B: [N][K] (row n); element size es bytes; D[m][n] = sum_k A[m][k] B[n][k]
static int run_gemm(const std::string& name, const Opts& o) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  const bool i8 = o.get("i8", 0) != 0;
  g.kind = i8 ? 2 : 0;
  const int es = i8 ? 1 : 2;
  const int KU = i8 ? 32 : 16;  // K per instruction
  g.N = (int)o.get("N", 256);
  g.ksteps = (int)o.get("ksteps", 4);
  g.repeat = (int)o.get("repeat", 1);
  const int K = KU * g.ksteps;
  const int a_mn = (int)o.get("a_mn", 0);       // A stored MN-major (image built by the host)
  g.a_src = (int)o.get("a_src", 0);
  g.b_src = (int)o.get("b_src", 0);
  const int a_swb = (int)o.get("a_swz", 128), b_swb = (int)o.get("b_swz", 128);  // swizzle span in bytes (0 = none)
  auto sw_code = [](int b) { return b == 128 ? SW_128B : b == 64 ? SW_64B : b == 32 ? SW_32B : SW_NONE; };
  g.a_sw = sw_code(a_swb);
  g.b_sw = sw_code(b_swb);
  const int rowA = K * es, rowB = K * es;  // bytes per K-major row (must equal the swizzle span when swizzled)
  g.a_bytes = a_mn ? 128u * K : 128u * rowA;
  g.b_bytes = (uint32_t)g.N * rowB;
  g.a_lbo = (uint32_t)o.get("a_lbo", 16);
  g.b_lbo = (uint32_t)o.get("b_lbo", 16);
  g.a_sbo = (uint32_t)o.get("a_sbo", a_mn ? 1024 : 8 * rowA);
  g.b_sbo = (uint32_t)o.get("b_sbo", 8 * rowB);
  g.a_kadv = (uint32_t)o.get("a_kadv", a_mn ? KU * 128 : KU * es);
  g.b_kadv = (uint32_t)o.get("b_kadv", KU * es);
  const int a_unsigned = (int)o.get("a_u8", 0);
  g.idesc = i8 ? instr_desc(D_S32, a_unsigned ? AB_U8 : AB_S8, AB_S8, a_mn, 0, 128, g.N)
               : instr_desc(D_F32, AB_F16, AB_F16, a_mn, 0, 128, g.N);
  const bool real = o.get("real", 0) != 0;  // non-integer fp16 data (accumulation-precision experiment)

  std::mt19937 rng(1234);
  std::vector<double> Ad((size_t)128 * K), Bd((size_t)g.N * K);
  std::vector<uint8_t> Araw((size_t)128 * K * es), Braw((size_t)g.N * K * es);
  auto put = [&](std::vector<uint8_t>& raw, std::vector<double>& dv, size_t idx, bool is_a) {
    if (i8) {
      int v = is_a ? (a_unsigned ? (int)(rng() % 256) : (int)(rng() % 256) - 128) : (int)(rng() % 14) - 3;
      raw[idx] = (uint8_t)v;
      dv[idx] = v;
    } else {
      float f = real ? 0.5f + (float)(rng() % 4096) / 8192.f : (float)((int)(rng() % 9) - 4);
      __half h = __float2half(f);
      memcpy(&raw[idx * 2], &h, 2);
      dv[idx] = (double)__half2float(h);
    }
  };
  for (size_t i = 0; i < Ad.size(); ++i) put(Araw, Ad, i, true);
  for (size_t i = 0; i < Bd.size(); ++i) put(Braw, Bd, i, false);

  // host-built shared-memory images
  std::vector<uint8_t> Aimg(g.a_src == 2 ? (size_t)128 * K * 2 : g.a_bytes), Bimg(g.b_bytes);
  if (g.a_src == 1) {
    if (a_mn) {  // MN-major, 128-byte swizzle: byte (m, k) at k*128 + ((m/16) ^ (k%8))*16 + m%16  (int8 only)
      for (int m = 0; m < 128; ++m)
        for (int k = 0; k < K; ++k) Aimg[(size_t)k * 128 + (((m / 16) ^ (a_swb == 128 ? (k & 7) : 0)) * 16) + m % 16] = Araw[(size_t)m * K + k];
    } else if (a_swb == 0) {  // no swizzle: 8-row x 16-byte core matrices, K direction at a_lbo, row groups at a_sbo
      for (int m = 0; m < 128; ++m)
        for (int j = 0; j < rowA; ++j)
          Aimg[(size_t)(m / 8) * g.a_sbo + (size_t)(j / 16) * g.a_lbo + (m % 8) * 16 + j % 16] = Araw[(size_t)m * rowA + j];
    } else {
      for (int m = 0; m < 128; ++m)
        for (int j = 0; j < rowA; ++j) Aimg[swz(a_swb, rowA, m, j)] = Araw[(size_t)m * rowA + j];
    }
  } else if (g.a_src == 2) {
    memcpy(Aimg.data(), Araw.data(), Araw.size());
    g.a_tmem_words = K / 2;
  }
  if (g.b_src == 1) {
    if (b_swb == 0) {
      for (int n = 0; n < g.N; ++n)
        for (int j = 0; j < rowB; ++j)
          Bimg[(size_t)(n / 8) * g.b_sbo + (size_t)(j / 16) * g.b_lbo + (n % 8) * 16 + j % 16] = Braw[(size_t)n * rowB + j];
    } else {
      for (int n = 0; n < g.N; ++n)
        for (int j = 0; j < rowB; ++j) Bimg[swz(b_swb, rowB, n, j)] = Braw[(size_t)n * rowB + j];
    }
  }
  uint8_t *dA, *dB, *dAimg, *dBimg;
  uint32_t* dD;
  CK(cudaMalloc(&dA, Araw.size()));
  CK(cudaMalloc(&dB, Braw.size()));
  CK(cudaMalloc(&dAimg, Aimg.size() + 16));
  CK(cudaMalloc(&dBimg, Bimg.size() + 16));
  CK(cudaMalloc(&dD, (size_t)128 * g.N * 4));
  CK(cudaMemcpy(dA, Araw.data(), Araw.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, Braw.data(), Braw.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dAimg, Aimg.data(), Aimg.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dBimg, Bimg.data(), Bimg.size(), cudaMemcpyHostToDevice));
  CUtensorMap mapA, mapB;
  memset(&mapA, 0, sizeof(mapA));
  memset(&mapB, 0, sizeof(mapB));
  auto tsw = [](int b) { return b == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : b == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : b == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE; };
  const CUtensorMapDataType dt = i8 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  if (g.a_src == 0 && !make_map_2d(&mapA, dA, dt, K, 128, (uint64_t)rowA, K, 128, tsw(a_swb))) { printf("%s: map A failed FAIL\n", name.c_str()); return 1; }
  if (g.b_src == 0 && !make_map_2d(&mapB, dB, dt, K, g.N, (uint64_t)rowB, K, g.N, tsw(b_swb))) { printf("%s: map B failed FAIL\n", name.c_str()); return 1; }
  const size_t smem = ((g.a_bytes + 1023) & ~1023u) + g.b_bytes + 2048;
  CK(cudaFuncSetAttribute(k_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  k_gemm<<<1, 128, smem>>>(mapA, mapB, g, dAimg, dBimg, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("%s: kernel error %s FAIL\n", name.c_str(), cudaGetErrorString(e));
    return 1;
  }
  std::vector<uint32_t> D((size_t)128 * g.N);
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  double max_err = 0, sum_rel = 0, max_rel = 0;
  int bad = 0, shown = 0;
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < g.N; ++n) {
      double ref = 0;
      for (int k = 0; k < K; ++k) ref += Ad[(size_t)m * K + k] * Bd[(size_t)n * K + k];
      ref *= g.repeat;
      double got;
      if (i8) got = (double)(int32_t)D[(size_t)m * g.N + n];
      else { float f; memcpy(&f, &D[(size_t)m * g.N + n], 4); got = f; }
      const double err = fabs(got - ref);
      max_err = fmax(max_err, err);
      if (real) {
        const double rel = (got - ref) / ref;
        sum_rel += rel;
        max_rel = fmax(max_rel, fabs(rel));
      } else if (err != 0) {
        ++bad;
        if (shown++ < 6) printf("   (m=%d,n=%d) got %.1f want %.1f\n", m, n, got, ref);
      }
    }
  if (real)
    printf("%s: repeat=%d mean_rel_err=%.3e max_rel_err=%.3e (2^-24 = 5.96e-8) INFO\n", name.c_str(), g.repeat,
           sum_rel / (128.0 * g.N), max_rel);
  else
    printf("%s: N=%d K=%d mismatches=%d max_err=%.1f %s\n", name.c_str(), g.N, K, bad, max_err, bad ? "FAIL" : "PASS");
  return bad != 0;
}

// ---------------------------------------------------------------------------------------------------------
// TMA delivery rate: every CTA (one per SM) keeps `depth` tile loads in flight from an L2-resident buffer.
//   mode 0: 2-D tensor map, box = rows x 128 B, source rows `pitch` bytes apart (strided, like XhT / c8)
//   mode 1: 1-D bulk copy of the same number of bytes from a contiguous tile
__global__ void __launch_bounds__(128) k_tma_bw(const __grid_constant__ CUtensorMap map, const uint8_t* __restrict__ src, int mode,
                                                int iters, uint32_t tile_bytes, int tiles_x, int tiles_y, int box_rows, int box_cols,
                                                int depth, int waitmode, unsigned long long* cycles) {
  extern __shared__ uint8_t bw_raw[];
  __shared__ uint64_t full[8];
  uint8_t* smem = bw_raw + ((1024u - (s32(bw_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < 8; ++s) bar_init(&full[s], 1);
    bar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    const int n_tiles = tiles_x * tiles_y;
    for (int it = 0; it < iters + depth; ++it) {
      const int s = it % depth;
      if (it >= depth) {
        const uint32_t par = ((it / depth) - 1) & 1, ba = s32(&full[s]);
        if (waitmode == 0) bar_wait(&full[s], par);
        else if (waitmode == 1) {  // non-blocking test_wait in a spin loop
          uint32_t done = 0;
          while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(ba), "r"(par) : "memory");
        } else {  // try_wait with a short suspend-time hint (ns)
          uint32_t done = 0;
          while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\nselp.u32 %0, 1, 0, p;\n}\n"
                         : "=r"(done) : "r"(ba), "r"(par), "r"(20u) : "memory");
        }
      }
      if (it < iters) {
        const int tile = (int)(((long long)it * gridDim.x + blockIdx.x) % n_tiles);
        bar_expect_tx(&full[s], tile_bytes);
        if (mode == 0) tma_load_2d(smem + (size_t)s * tile_bytes, &map, &full[s], (tile % tiles_x) * box_cols, (tile / tiles_x) * box_rows);
        else if (mode >= 2) {  // the tile as `mode` linear pieces on the same barrier
          const uint32_t piece = tile_bytes / mode;
          for (int q = 0; q < mode; ++q)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             s32(smem + (size_t)s * tile_bytes + q * piece)),
                         "l"(src + (size_t)tile * tile_bytes + q * piece), "r"(piece), "r"(s32(&full[s]))
                         : "memory");
        } else
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                           s32(smem + (size_t)s * tile_bytes)),
                       "l"(src + (size_t)tile * tile_bytes), "r"(tile_bytes), "r"(s32(&full[s]))
                       : "memory");
      }
    }
    cycles[blockIdx.x] = (unsigned long long)(clock64() - t0);
  }
}

static int run_tma_bw(int mode, int box_rows, int row_bytes, long pitch, int depth, int swz_b, int waitmode) {
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const uint32_t tile_bytes = (uint32_t)box_rows * row_bytes;
  // a 64 MB (L2-resident) source: rows of `pitch` bytes
  const size_t total = (size_t)64 << 20;
  const long rows = (long)(total / pitch);
  const int tiles_x = (int)(pitch / row_bytes), tiles_y = (int)(rows / box_rows);
  uint8_t* src;
  CK(cudaMalloc(&src, total + 4096));
  CK(cudaMemset(src, 1, total));
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, sms * 8));
  CUtensorMap map;
  memset(&map, 0, sizeof(map));
  CUtensorMapSwizzle sw = swz_b == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : swz_b == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                          : swz_b == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  if (mode == 0 && !make_map_2d(&map, src, CU_TENSOR_MAP_DATA_TYPE_UINT8, (uint64_t)pitch, (uint64_t)rows, (uint64_t)pitch,
                                (uint32_t)row_bytes, (uint32_t)box_rows, sw)) {
    printf("tma_bw: map failed FAIL\n");
    return 1;
  }
  const int iters = 2000;
  const size_t smem = (size_t)depth * tile_bytes + 2048;
  CK(cudaFuncSetAttribute(k_tma_bw, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  k_tma_bw<<<sms, 128, smem>>>(map, src, mode, 200, tile_bytes, tiles_x, tiles_y, box_rows, row_bytes, depth, waitmode, cyc);
  CK(cudaEventRecord(e0));
  k_tma_bw<<<sms, 128, smem>>>(map, src, mode, iters, tile_bytes, tiles_x, tiles_y, box_rows, row_bytes, depth, waitmode, cyc);
  CK(cudaEventRecord(e1));
  CK(cudaDeviceSynchronize());
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<unsigned long long> h(sms);
  CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
  double mean = 0;
  for (auto v : h) mean += (double)v / sms;
  const double bytes = (double)sms * iters * tile_bytes;
  printf("tma_bw mode=%d wait=%d box=%dx%dB pitch=%ld depth=%d swz=%d: %.2f TB/s, %.1f B/clk/SM, %.0f clk per tile (%u B) INFO\n", mode, waitmode, box_rows,
         row_bytes, pitch, depth, swz_b, bytes / (ms * 1e-3) / 1e12, (double)iters * tile_bytes / mean, mean / iters, tile_bytes);
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) return 1;
  const std::string cmd = argv[1];
  if (cmd == "tmem") return run_tmem();
  if (cmd == "dump" && argc >= 6) return run_dump(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atoi(argv[5]));
  if (cmd == "tma_bw" && argc >= 8)
    return run_tma_bw(atoi(argv[2]), atoi(argv[3]), atoi(argv[4]), atol(argv[5]), atoi(argv[6]), atoi(argv[7]), argc > 8 ? atoi(argv[8]) : 0);
  if (cmd == "gemm" && argc >= 3) {
    Opts o;
    for (int i = 3; i < argc; ++i) {
      const char* eq = strchr(argv[i], '=');
      if (eq) o.kv[std::string(argv[i], eq - argv[i])] = atol(eq + 1);
    }
    return run_gemm(argv[2], o);
  }
  printf("unknown experiment\n");
  return 1;
}
