// Hand-written sm_100a PTX wrappers: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05
// (alloc / mma.block_scale / cp / ld / commit), cluster helpers and UMMA descriptors.
// No CUTLASS / CuTe: the bit layouts below are the hardware contracts, cross-checked against
// the (read-only) notes in SURVEY.md Appendix B.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b200q {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "elect.sync _|P, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ cluster
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_sync() {
  cluster_arrive();
  cluster_wait();
}
// map a local shared::cta address to the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

// ------------------------------------------------------------------ programmatic dependent launch
// wait: block until the prerequisite grid (the previous kernel in the stream) has completed and its writes are
// visible; launch_dependents: allow the next kernel in the stream (if launched with the PDL attribute) to start.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive on a barrier living in another CTA of the cluster (address from mapa).  Default semantics (.release at CTA
// scope), NOT .release.cluster: the cluster-scope form compiles to MEMBAR.ALL.GPU + ERRBAR in front of the arrive, a
// round trip to L2 on the accumulator hand-off path (ncu: the most-sampled stall of the epilogue warps).  What the
// arrive publishes here lives in TMEM / registers and is ordered by tcgen05.wait::ld + tcgen05.fence::before_thread_sync.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
// Asynchronous stores into the shared memory of a CTA of the cluster whose completion is counted (in bytes) on an mbarrier of
// THAT CTA: the producer needs no fence and no arrive -- the consumer's wait on the barrier (expect_tx armed once) orders the data.
__device__ __forceinline__ void st_async_v4(uint32_t cluster_addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
               "r"(a), "r"(b), "r"(c), "r"(d), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t cluster_addr, uint32_t v, uint32_t cluster_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(cluster_addr), "r"(v), "r"(cluster_bar)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

#ifndef B200Q_SPIN_LIMIT_CYCLES
#define B200Q_SPIN_LIMIT_CYCLES 4000000000ll  // ~2 s: a lost arrival traps instead of hanging the GPU
#endif

template <bool kCluster = false>
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int tag = 0) {
  if (kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!(kCluster ? mbar_try_wait_cluster(bar, parity) : mbar_try_wait(bar, parity))) {
    if (clock64() - t0 > B200Q_SPIN_LIMIT_CYCLES) {
      printf("b200q: mbarrier wait timed out (tag %d, block %d, thread %d, parity %u)\n", tag, (int)blockIdx.x,
             (int)threadIdx.x, parity);
      __trap();
    }
  }
}

// Bare wait for the hot single-thread issue loops: no time-out / printf path (whose call keeps the loop state out of the
// uniform registers).  Only where a lost arrival is impossible by construction or already caught by a mbar_wait elsewhere.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%0], %1;\n"
      "@!P bra WAIT_%=;\n"
      "}\n" ::"r"(bar), "r"(parity)
      : "memory");
}

// ------------------------------------------------------------------ cross-CTA progress counters in global memory
// producer side: all writes of the warp, __syncwarp, then ONE lane publishes with a gpu-scope release add;
// consumer side: acquire load, then a generic->async proxy fence so that a following TMA load sees the data.
__device__ __forceinline__ void red_release_gpu_add(uint32_t* addr, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ void wait_counter_ge(const uint32_t* addr, uint32_t want, int tag = 0) {
  if (ld_acquire_gpu(addr) >= want) return;
  const long long t0 = clock64();
  while (ld_acquire_gpu(addr) < want) {
    __nanosleep(100);
    if (clock64() - t0 > B200Q_SPIN_LIMIT_CYCLES) {
      printf("b200q: progress-counter wait timed out (tag %d, block %d, thread %d, want %u)\n", tag, (int)blockIdx.x,
             (int)threadIdx.x, want);
      __trap();
    }
  }
}

// ------------------------------------------------------------------ TMA
__device__ __forceinline__ void prefetch_tensormap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
// 2-D / 3-D tiled loads global -> this CTA's shared memory.  kCtaGroup == 2: the completion may be signalled on
// the mbarrier of EITHER CTA of the pair (bar is then a shared::cluster address, e.g. mapa(bar, 0) = the leader's).
template <int kCtaGroup = 1>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1) {
  if constexpr (kCtaGroup == 1)
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
// 2-D tiled load multicast to the CTAs of `cta_mask` (same shared-memory offset in each).  kCtaGroup == 2: each
// destination's completion is signalled on the barrier of ITS pair leader -- `bar` is the issuing CTA's own pair-leader
// barrier as a shared::cluster address (the convention of CUTLASS' SM100_TMA_2SM_LOAD_MULTICAST: peer bit cleared).
template <int kCtaGroup = 2>
__device__ __forceinline__ void tma_load_2d_multicast(uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1,
                                                      uint16_t cta_mask) {
  static_assert(kCtaGroup == 2, "only the CTA-pair form is used");
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
template <int kCtaGroup = 1>
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int32_t c0, int32_t c1,
                                            int32_t c2) {
  if constexpr (kCtaGroup == 1)
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  else
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// L2 prefetch of a tensor-map box (no shared-memory destination, no barrier).  Always safe with respect to later writers of
// that memory: L2 is the coherence point, a line written afterwards is simply updated in place.
__device__ __forceinline__ void tma_prefetch_2d(const void* tmap, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(tmap), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int32_t c0, int32_t c1, int32_t c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
// 1-D bulk copy global -> shared (size multiple of 16, both 16-byte aligned)
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
// 2-D tiled store shared -> global
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t src, int32_t c0, int32_t c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------ tcgen05: TMEM management
template <int kCtaGroup>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
}
template <int kCtaGroup>
__device__ __forceinline__ void tmem_relinquish() {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  else
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int kCtaGroup>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ tcgen05: descriptors
// 64-bit shared-memory matrix descriptor (K-major operand):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4 (distance between 8-row groups)
//   [46,48) version = 1 on sm_100     [49,52) base offset   [61,64) layout: 0 none, 2 = 128B, 4 = 64B, 6 = 32B swizzle
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout_type & 7u) << 61;
  return d;
}
constexpr uint32_t kLayoutNone = 0, kLayoutSw128 = 2, kLayoutSw64 = 4, kLayoutSw32 = 6;

// 32-bit instruction descriptor for kind::mxf4 / kind::mxf4nvf4 block-scaled MMA:
//   [4,6) b_sf_id  [7,10) a_format (E2M1 = 1)  [10,13) b_format (E2M1 = 1)  15 a_major (0 = K)  16 b_major (0 = K)
//   [17,23) N >> 3   23 scale format (0 = UE4M3, 1 = UE8M0)   [24,29) M >> 4   [29,31) a_sf_id   31 k_size (0 = K64)
__host__ __device__ constexpr uint32_t make_idesc_fp4(int M, int N, bool ue8m0) {
  return (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((ue8m0 ? 1u : 0u) << 23) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t idesc_with_sf_id(uint32_t idesc, uint32_t a_sf_id, uint32_t b_sf_id) {
  return idesc | (b_sf_id << 4) | (a_sf_id << 29);
}

// ------------------------------------------------------------------ tcgen05: MMA / cp / commit / ld
// D[tmem] (+)= A[smem] * B[smem] with per-block scales from TMEM.  One thread issues.
//   kNV: kind::mxf4nvf4 block16 (ue4m3 scales);  kF8: kind::mxf8f6f4 block32 (8-bit operands, K = 32);
//   otherwise kind::mxf4 block32 (ue8m0 scales, K = 64).
#define B200Q_MMA_BS(CG, KINDSTR)                                                                                  \
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"                                                        \
               "tcgen05.mma.cta_group::" CG ".kind::" KINDSTR " [%0], %1, %2, %3, [%5], [%6], p;\n}\n" ::"r"(tmem_d), \
               "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)                   \
               : "memory")
template <int kCtaGroup, bool kNV, bool kF8 = false>
__device__ __forceinline__ void mma_fp4_block_scaled(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                     uint32_t tmem_sfa, uint32_t tmem_sfb, uint32_t accumulate) {
  if constexpr (kCtaGroup == 1) {
    if constexpr (kF8) B200Q_MMA_BS("1", "mxf8f6f4.block_scale");
    else if constexpr (kNV) B200Q_MMA_BS("1", "mxf4nvf4.block_scale.block16");
    else B200Q_MMA_BS("1", "mxf4.block_scale.block32");
  } else {
    if constexpr (kF8) B200Q_MMA_BS("2", "mxf8f6f4.block_scale");
    else if constexpr (kNV) B200Q_MMA_BS("2", "mxf4nvf4.block_scale.block16");
    else B200Q_MMA_BS("2", "mxf4.block_scale.block32");
  }
}
#undef B200Q_MMA_BS

// bf16 x bf16 -> fp32 (kind::f16), used by the probe / rotation experiments
template <int kCtaGroup>
__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (kCtaGroup == 1)
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// 32 rows x 128 bit shared -> TMEM, replicated to the 4 lane quarters (scale-factor copy)
template <int kCtaGroup>
__device__ __forceinline__ void tmem_cp_32x128b_warpx4(uint32_t tmem_dst, uint64_t smem_desc) {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
  else
    asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(tmem_dst), "l"(smem_desc) : "memory");
}

// all previously issued tcgen05 async ops of this thread arrive on `bar` when complete
template <int kCtaGroup>
__device__ __forceinline__ void tc_commit(uint32_t bar, uint16_t cta_mask = 3) {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(cta_mask)
                 : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x4(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_32x32b_x1(uint32_t taddr, uint32_t v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace ptx
}  // namespace b200q
