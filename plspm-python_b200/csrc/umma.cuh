// tcgen05 / TMEM / TMA building blocks (sm_100a inline PTX) shared by the tensor-core kernels
// (kernels_vote_mma.cuh, kernels_gram_mma.cuh) and by tools/umma_probe.cu, which validates every
// piece of this file against a CPU reference on the GPU box.  No CUTLASS: descriptors are assembled
// by hand; the bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptor" tables.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace umma {

// ---- shared-memory matrix descriptor (64 bit) --------------------------------------------------------
//  [0,14)  start address >> 4        [16,30) leading-dimension byte offset >> 4
//  [32,46) stride-dimension byte offset >> 4     [46,48) version = 1 (sm_100)
//  [49,52) base offset (0: tiles are aligned to the swizzle period)   [61,64) layout / swizzle mode
enum Swizzle : uint32_t { SW_NONE = 0, SW_128B = 2, SW_64B = 4, SW_32B = 6 };
__host__ __device__ inline uint64_t smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t swizzle) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(swizzle & 7u) << 61;
  return d;
}
// advancing along K inside a swizzle atom (K-major) or to the next group of k-rows (MN-major): the start
// address field moves by the byte offset of the un-swizzled position
__host__ __device__ inline uint64_t desc_advance(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// ---- instruction descriptor (32 bit) -----------------------------------------------------------------
//  [4,6) D format (0 f16, 1 f32, 2 s32)   [7,10) A format   [10,13) B format
//  (kind::f16: 0 f16, 1 bf16; kind::tf32: 2; kind::i8: 0 u8, 1 s8)
//  [15] A major (0 K, 1 MN)   [16] B major   [17,23) N >> 3   [24,29) M >> 4
enum DFmt : uint32_t { D_F16 = 0, D_F32 = 1, D_S32 = 2 };
enum ABFmt : uint32_t { AB_F16 = 0, AB_BF16 = 1, AB_TF32 = 2, AB_U8 = 0, AB_S8 = 1 };
__host__ __device__ inline uint32_t instr_desc(uint32_t dfmt, uint32_t afmt, uint32_t bfmt, uint32_t a_mn_major,
                                               uint32_t b_mn_major, uint32_t M, uint32_t N) {
  return (dfmt << 4) | (afmt << 7) | (bfmt << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((N >> 3) << 17) |
         ((M >> 4) << 24);
}

#if defined(__CUDACC__)
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(bar)), "r"(count));
}
__device__ __forceinline__ void bar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void bar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(bar)) : "memory");
}
__device__ __forceinline__ bool bar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(s32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
// A wait that cannot hang the device: a protocol error (a barrier that never completes) traps after ~2 s with the
// barrier's tag instead of spinning forever, so the host sees a launch failure and the tests a clear message.
__device__ __noinline__ void bar_timeout(int tag, uint32_t parity) {
  printf("plspm_b200: mbarrier wait timed out (tag %d, parity %u, block %d, thread %d)\n", tag, parity, (int)blockIdx.x,
         (int)threadIdx.x);
  __trap();
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (bar_try(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!bar_try(bar, parity)) {
    if ((++spins & 0xfffu) == 0 && clock64() - t0 > 4000000000ll) bar_timeout(tag, parity);
  }
}

// ---- the same primitives on precomputed 32-bit shared addresses -------------------------------------------
// The issuing threads of the tensor-core kernels are single threads running long dependent instruction chains; ncu
// showed them ISSUE-bound (no dominant wait, ~300 instructions per 64-row chunk): the compiler re-derives the
// cluster-window address of every __shared__ object at each use (S2R SR_CgaCtaId + LEA) and inlines the watchdog
// into every wait.  Hot loops therefore take their barrier / tile addresses once (s32()) and use these.
__device__ __forceinline__ bool bar_try_a(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// (returns the cycles spent waiting: pipeline diagnostics, see PLSPM_KERNEL_STATS)
__device__ __noinline__ long long bar_wait_slow_a(uint32_t bar, uint32_t parity, int tag) {
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!bar_try_a(bar, parity)) {
    if ((++spins & 0xfffu) == 0 && clock64() - t0 > 4000000000ll) bar_timeout(tag, parity);
  }
  return clock64() - t0;
}
__device__ __forceinline__ long long bar_wait_a(uint32_t bar, uint32_t parity, int tag = 0) {
  if (!bar_try_a(bar, parity)) return bar_wait_slow_a(bar, parity, tag);
  return 0;
}
__device__ __forceinline__ void bar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void bar_expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_2d_mc_a(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// linear bulk copy global -> shared (TMA engine, no tensor map): the operands of the tensor-core kernels are stored
// in HBM as ready-made shared-memory tile images, so a whole tile is ONE request.  Measured (tools/umma_probe tma_bw):
// 2-D boxes of 128-byte rows deliver 11 TB/s chip-wide whatever the pitch; linear copies 13 TB/s at 32 KB, 20 TB/s at 64 KB.
__device__ __forceinline__ void bulk_load_a(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
               "r"(bar)
               : "memory");
}
__device__ __forceinline__ void mma_commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mma_commit_mc_a(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}

// ---- TMA (cp.async.bulk.tensor), 2-D tiles -------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
// coordinates: c0 = innermost (contiguous) dimension, c1 = outer dimension, in elements
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          s32(dst_smem)),
      "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// the same tile delivered to the same shared-memory offset of every CTA of the cluster named in `mask`; each
// destination's mbarrier (same offset) receives the bytes
__device__ __forceinline__ void tma_load_2d_mc(void* dst_smem, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
          s32(dst_smem)),
      "l"(map), "r"(s32(bar)), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
// ---- thread-block clusters ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// generic-proxy writes (st.shared) that the tensor core / TMA (async proxy) will read
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- TMEM ----------------------------------------------------------------------------------------------
// one warp allocates (power of two >= 32 columns); the base address lands in shared memory
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// all previously issued tcgen05.mma of this thread done -> one arrival on the mbarrier
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(bar)) : "memory");
}

// ... one arrival on the mbarrier at this offset in EVERY CTA of the cluster named in `mask`
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s32(bar)),
               "h"(mask)
               : "memory");
}

// ---- tcgen05.mma (issued by ONE thread) ----------------------------------------------------------------
// D[tmem] (+)= A[smem desc] * B[smem desc]^T ; accumulate = 0 overwrites D
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// A from tensor memory (lane = row m, 32-bit column = two consecutive k, low half first)
__device__ __forceinline__ void mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_i8_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- TMEM <-> registers: shape 32x32b (thread t of the warp owns lane 32*(warp%4) + t, N consecutive columns)
#define UMMA_R8(r, o) "=r"(r[o + 0]), "=r"(r[o + 1]), "=r"(r[o + 2]), "=r"(r[o + 3]), "=r"(r[o + 4]), "=r"(r[o + 5]), "=r"(r[o + 6]), "=r"(r[o + 7])
#define UMMA_W8(r, o) "r"(r[o + 0]), "r"(r[o + 1]), "r"(r[o + 2]), "r"(r[o + 3]), "r"(r[o + 4]), "r"(r[o + 5]), "r"(r[o + 6]), "r"(r[o + 7])
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : UMMA_R8(r, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : UMMA_R8(r, 0), UMMA_R8(r, 8)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, "
      "[%32];"
      : UMMA_R8(r, 0), UMMA_R8(r, 8), UMMA_R8(r, 16), UMMA_R8(r, 24)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::UMMA_W8(r, 0), "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};" ::UMMA_W8(r, 0),
      UMMA_W8(r, 8), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};" ::
          UMMA_W8(r, 0),
      UMMA_W8(r, 8), UMMA_W8(r, 16), UMMA_W8(r, 24), "r"(taddr)
      : "memory");
}
// TMEM address of (lane, column) relative to an allocation base
__device__ __forceinline__ uint32_t tmem_addr(uint32_t base, uint32_t lane, uint32_t col) { return base + (lane << 16) + col; }
#endif  // __CUDACC__

// ---- host: tensor maps through the driver entry point (no link-time dependency on libcuda) --------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// 2-D row-major tensor [rows][cols] of `esize`-byte elements with a row pitch of `pitch_bytes`; box = box_cols x box_rows.
// Out-of-bounds elements of a box read as zero.
inline bool make_map_2d(CUtensorMap* map, const void* base, CUtensorMapDataType dt, uint64_t cols, uint64_t rows,
                        uint64_t pitch_bytes, uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle sw) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {pitch_bytes};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace umma
