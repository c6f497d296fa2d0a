// Block-scaled FP4 x FP4 -> bf16 GEMM for sm_100a, hand-written tcgen05 / TMEM / TMA.
//
// Replaces the reference's CUTLASS CollectiveBuilder kernels
// (qutlass/csrc/gemm.cu:174-326: matmul_host_mxf4_bf16_tn / matmul_host_nvf4_bf16_tn).
//
//   D[M,N] = bf16_rne( alpha * sum_k (A[m,k] * SFA[m,k/g]) * (B[n,k] * SFB[n,k/g]) ),  fp32 accumulation in TMEM
//   A [M,K/2], B [N,K/2] packed e2m1 (K-major); SFA/SFB in the cuBLAS block-scaled layout
//   (512-B blocks = 128 rows x 4 scales, K-blocks fastest), g = 32 (ue8m0) or 16 (ue4m3).
//
// Kernel shape (persistent, warp-specialised, 320 threads, 1 CTA / SM):
//   warp 0     TMA producer (one lane): per k-tile (256 K-elements = one 128-byte swizzle row) FOUR tensor-map
//              loads -- A tile (128 rows), this CTA's share of the B tile, SFA blocks, SFB blocks -- into a
//              shared-memory stage; completion by mbarrier complete_tx.  With kCtaGroup == 2 both CTAs of the
//              pair signal the LEADER's barrier (cp.async.bulk.tensor ... .cta_group::2).
//   warp 1     MMA issuer (one lane, leader CTA only): scale blocks smem -> TMEM with tcgen05.cp
//              (32x128b.warpx4), then 4 x tcgen05.mma.kind::mxf4[nvf4].block_scale (K = 64 each);
//              tcgen05.commit frees the stage (multicast to both CTAs) and, after the last k-tile, publishes
//              the accumulator.  With kCtaGroup == 2 the pair computes a 256 x BN tile, each CTA holding 128
//              rows of A, BN/2 rows of B and its 128 x BN fp32 accumulator.
//   warps 2-9  epilogue (8 warps = 4 TMEM lane quarters x 2 column halves): drain the WHOLE accumulator half
//              into registers with back-to-back tcgen05.ld (<= 128 registers / thread) and immediately release
//              TMEM -- so even the single-buffered 256-column accumulator only stalls the tensor pipe for the
//              drain, not for the stores -- then alpha (device scalar, fp32), RNE to bf16, 128B-swizzled
//              shared-memory staging and TMA tensor stores (full 128-byte lines; clips the M/N tails).
//
// Bit-exactness (reference tests demand out == bf16(fp64 matmul)): a single fp32 accumulation chain
// per output over all of K (no split-K), alpha applied once in fp32, one RNE to bf16.
#include "common.cuh"
#include "ptx.cuh"

// Profiling switches (timing-only ablations, timelines) exist only in the -DB200Q_PROFILING build: in the product build the
// flag word is the constant 0 and the compiler removes every branch on it -- also from the single-thread issue loops, where a
// dormant time-out wait behind a run-time flag cost the small-M kernels 1.5 us (profiles/r02_notes.md).
#ifndef B200Q_FLAGS
#ifdef B200Q_PROFILING
#define B200Q_FLAGS(p) ((p).flags)
#else
#define B200Q_FLAGS(p) 0
#endif
#endif
#include "quantize_tile.cuh"
#include "tmap.cuh"
#include <type_traits>

#include <cuda.h>
#include <mutex>
#include <stdlib.h>

namespace b200q {
using namespace ptx;

constexpr int BM = 128;          // rows of A per CTA
constexpr int BK_BYTES = 128;    // one k-tile = 256 e2m1 = 128 bytes per row (one 128B-swizzle row)
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;   // 320
// fused kernel: kFuse (2 or 4) extra warps rotate + quantise the activations.  4 warps -> 448 threads -> 4 warps on one
// scheduler -> 128 registers / thread (chunked epilogue drain); 2 warps -> 384 threads -> 168 registers (plain epilogue).
constexpr int kSmemBudget = 227 * 1024;

// kF8: 0 = FP4 operands; 1 = MXFP8 "tn" (A [M, K] bytes); 2 = MXFP8 "nn" (A stored [K, M], M contiguous: an MN-major
// tcgen05 operand -- same 128B-swizzled 16 KB stage, rows are K instead of M)
template <int kCtaGroup, int BN, bool kNV, int A_ROWS = 128, int kF8 = 0, int kFuse = 0>
struct GemmCfg {
  // A_ROWS < 128 (small M): only A_ROWS rows of the A tile are loaded and kept per stage; the MMA still reads a
  // 128-row operand (the bytes that follow) -- those accumulator rows are garbage and never stored.  Smaller
  // stages = more k-tiles of B in flight, which is what bounds the weight-streaming (decode) regime.
  static_assert(A_ROWS % 8 == 0 && A_ROWS >= 8 && A_ROWS <= 128 && (A_ROWS == 128 || kCtaGroup == 1), "A_ROWS");
  static_assert(kF8 != 2 || A_ROWS == 128, "the MN-major A tile is always 128 K-rows x 128 M-bytes");
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "BN must be a multiple of 64 in [64, 256]");
  // kF8: 8-bit e4m3 operands (MXFP8, ue8m0 scales per 32): a 128-byte k-tile is 128 elements = 4 scales = one block
  static constexpr int SFKB = kNV ? 4 : (kF8 ? 1 : 2);           // 512-B scale blocks per 128 rows per k-tile
  static constexpr int BK_ELEMS = kF8 ? 128 : 256;               // K elements per k-tile (128 bytes per row)
  static constexpr int MMA_K = kF8 ? 32 : 64;                    // K elements per tcgen05.mma (32 bytes per row)
  // a BN-wide tile starts at a multiple of 64 rows of B: its scales begin 0 or 2 TMEM columns into a block
  static constexpr int NB = (BN % 128 == 0) ? BN / 128 : (BN + 64 + 127) / 128;   // SFB row-blocks a tile can touch
  static constexpr int B_ROWS = BN / kCtaGroup;                  // B rows this CTA stages
  static constexpr int A_BYTES = A_ROWS * BK_BYTES;
  static constexpr int B_BYTES = B_ROWS * BK_BYTES;
  static constexpr int SFA_BYTES = SFKB * 512;
  static constexpr int SFB_BYTES = NB * SFKB * 512;
  static constexpr int LOAD_BYTES = A_BYTES + B_BYTES + SFA_BYTES + SFB_BYTES;      // bytes one CTA's TMA loads deliver per stage
  static constexpr int STAGE_BYTES = (LOAD_BYTES + 1023) / 1024 * 1024;            // stage pitch keeps the 128B-swizzle alignment
  static constexpr int SFA_COLS = SFKB * 4;
  static constexpr int SFB_COLS = SFKB * 4 * NB;
  static constexpr int SF_COLS = SFA_COLS + SFB_COLS;
  static constexpr int ACC_STAGES = (2 * BN + SF_COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_USED = ACC_STAGES * BN + SF_COLS;
  static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
  // epilogue: each of the 8 warps owns 32 rows x EPI_COLS columns, stored in chunks of EPI_CHUNK columns
  static constexpr int EPI_COLS = BN / 2;
  static constexpr int EPI_CHUNK = (EPI_COLS % 64 == 0) ? 64 : 32;
  static constexpr int EPI_NCHUNK = EPI_COLS / EPI_CHUNK;
  static constexpr int STG_BYTES = 32 * EPI_CHUNK * 2;           // one staging buffer per epilogue warp
  static constexpr int STG_TOTAL = kEpiWarps * STG_BYTES;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int QSTG_BYTES = kFuse * 2048;                // quantiser warps' swizzled staging (2 KB each)
  static constexpr int THREADS = kGemmThreads + 32 * kFuse;
  static constexpr int STAGES_RAW = (kSmemBudget - BAR_BYTES - 1024 - STG_TOTAL - QSTG_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 12 ? 12 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STG_TOTAL + BAR_BYTES + QSTG_BYTES + 1024;  // +1024 alignment slack
  static constexpr uint32_t TX_BYTES = (uint32_t)LOAD_BYTES * kCtaGroup;   // what the (leader's) full barrier expects
  static_assert(TMEM_USED <= 512, "TMEM overflow");
  static_assert(STAGES >= 2, "not enough shared memory for 2 stages");
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024-byte alignment for the 128B swizzle");
  static_assert(EPI_COLS % 32 == 0 && EPI_COLS <= 128, "epilogue register budget");
};

// Debug timeline (profiling flag 1 << 24): CTA 0 records clock64 per tile -- [0] MMA warp owns the accumulator,
// [1] first k-tile landed, [2] last MMA issued, [3] epilogue warp 0 sees the accumulator complete, [4] drained,
// [5] its stores issued.  Read back with b200q_debug_read_trace.
constexpr int kTraceTiles = 32, kTraceEvents = 8;
__device__ unsigned long long g_gemm_trace[kTraceTiles * kTraceEvents];
__device__ __forceinline__ void trace_event(int flags, int tile_idx, int ev) {
  if ((flags & (1 << 24)) && blockIdx.x == 0 && tile_idx < kTraceTiles && (threadIdx.x & 31) == 0)
    g_gemm_trace[tile_idx * kTraceEvents + ev] = clock64();
}
// Kernel-level events of CTA 0 (same flag): [0] clock64 at entry, [1] globaltimer (ns) at entry, [2] set-up done,
// [3] producer past griddepcontrol.wait, [5] clock64 at exit, [6] globaltimer at exit.  b200q_debug_read_ktrace.
__device__ unsigned long long g_gemm_ktrace[8];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void ktrace(int flags, int ev, bool wall = false) {
  if ((flags & (1 << 24)) && blockIdx.x == 0) g_gemm_ktrace[ev] = wall ? globaltimer_ns() : (unsigned long long)clock64();
}

struct GemmParams {
  const float* alpha;
  __nv_bfloat16* d;
  int M, N, K;
  int ldd;            // row pitch of D in elements (== N unless this launch covers a column slice of a wider D)
  int tiles_m;        // ceil(M / (BM * cta_group))   (cluster tiles along M)
  int tiles_n;        // ceil(N / BN)
  int k_tiles;        // ceil(K / 256)
  int tma_store;      // 1: TMA-store epilogue (needs N % 8 == 0); 0: direct global stores
  int flags;          // profiling builds only (-DB200Q_PROFILING): bit0 skip stores, bit1 skip TMEM loads, ...; else 0
  int static_weights; // 1: the caller guarantees B / SFB are not produced by the preceding kernels of the stream
                      //    (B200Q_GEMM_STATIC_WEIGHTS): their first ring of loads may be issued before griddepcontrol.wait
};

// Fused quantise + GEMM (kFuse): the activations are rotated + quantised by 4 extra warps of the SAME persistent kernel
// while its tensor pipe works, instead of by a separate kernel in front of it.
//   * quantiser warps take warp-tiles (1024 consecutive elements of x) from a global ticket counter in ascending
//     order (dynamic, so progress never depends on a CTA that is not resident yet), write e2m1 codes + scales to
//     global memory exactly like the standalone kernel, and bump ready[tile's 256-row block] with a release add;
//   * output tiles are walked N-fastest, so the first wave needs only the first one or two row blocks of A; the TMA
//     producer acquires ready[tm] == (rows in block) * K/1024 before the first A / SFA load of a tile of row block tm;
//   * the last CTA to finish zeroes the counters again (the workspace is zero on entry AND on exit).
struct FuseParams {
  QuantParams q;
  uint32_t* ctr;                 // [0] ticket, [1] finished CTAs, [2 + tm] warp-tiles of row block tm written
  uint32_t tiles_per_row;        // K / 1024
  int had, method;
};

// One warp's share of the activation quantisation.  Work is claimed in batches of BATCH consecutive warp-tiles from the
// ticket counter.  HELPER = false (dedicated quantiser warps): the next batch's ticket and the next tile's loads are
// always in flight under the current tile's math.  HELPER = true (epilogue warps before their first accumulator): no
// look-ahead -- a claimed batch must be finished, so the warp re-checks `stop_bar` (its CTA's first tmem-full
// barrier, phase 0) before every claim and leaves at most one batch late.
template <int HAD, bool NV, int METHOD, bool HELPER>
__device__ __forceinline__ void quantiser_loop(const FuseParams& fp, uint4* stage, int lane, uint32_t stop_bar) {
  constexpr uint32_t BATCH = HELPER ? 2u : 4u;
  const QuantParams& p = fp.q;
  const float c_scale = __bfloat162float(p.rot[0]);
  float gs = 1.f, gs_rcp = 1.f;
  if constexpr (NV) {
    gs = *p.gs;
    gs_rcp = rcp_approx_ftz(gs);
  }
  const uint32_t n_tiles = (uint32_t)p.n_tiles;
  const uint32_t tiles_per_block = 256u * fp.tiles_per_row;     // one cluster tile = 256 rows of A
  uint32_t* ready = fp.ctr + 2;
  auto claim = [&]() -> uint32_t {          // lane 0's register only; broadcast (and wait for it) with bcast()
    uint32_t t = 0;
    if (lane == 0) t = atomicAdd(fp.ctr, BATCH);
    return t;
  };
  auto bcast = [&](uint32_t t) -> uint32_t { return __shfl_sync(0xffffffffu, t, 0); };
  auto load = [&](uint32_t tile, uint4 (&dst)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dst[i] = (tile < n_tiles) ? __ldg(p.x + ((int64_t)tile * 128 + i * 32 + lane)) : make_uint4(0, 0, 0, 0);
  };
  auto publish = [&](uint32_t t0, uint32_t t1) {   // warp-tiles [t0, t1) are written: bump their row blocks' counters
    __syncwarp();                                  // every lane's stores precede lane 0's release
    if (lane == 0) {
      const uint32_t b0 = t0 / tiles_per_block, b1 = (t1 - 1u) / tiles_per_block;
      if (b0 == b1) {
        red_release_gpu_add(ready + b0, t1 - t0);
      } else {                                     // a batch can straddle one block boundary
        const uint32_t cut = b1 * tiles_per_block;
        __threadfence();
        atomicAdd(ready + b0, cut - t0);
        atomicAdd(ready + b1, t1 - cut);
      }
    }
  };
  uint32_t base = bcast(claim());
  uint32_t next_ticket = 0;
  uint4 nxt[4];
  load(base, nxt);
  while (base < n_tiles) {
    if constexpr (!HELPER) next_ticket = claim();            // result is not needed before this batch's last tile
    const uint32_t end = (base + BATCH < n_tiles) ? base + BATCH : n_tiles;
    uint32_t nbase = n_tiles;
    for (uint32_t tile = base; tile < end; ++tile) {
      uint4 ld[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) ld[i] = nxt[i];
      uint32_t ntile = tile + 1u;
      if (ntile == end) {
        if constexpr (!HELPER) { nbase = bcast(next_ticket); ntile = nbase; } else { ntile = n_tiles; }
      }
      load(ntile, nxt);
      float v[32];
      tile_stage_unpack(ld, stage, lane, v);
      tile_rotate_hadamard<HAD>(v, c_scale);
      tile_quantise_store<NV, METHOD, false>(p, v, (int64_t)tile, lane, gs, gs_rcp);
    }
    publish(base, end);
    if constexpr (HELPER) {
      if (bcast(lane == 0 ? (uint32_t)mbar_try_wait(stop_bar, 0) : 0u)) break;   // first accumulator ready: back to epilogue duty
      nbase = bcast(claim());
      load(nbase, nxt);
    }
    base = nbase;
  }
}

template <bool NV, bool HELPER>
__device__ __forceinline__ void quantiser_role(const FuseParams& fp, uint4* stage, int lane, uint32_t stop_bar) {
#define B200Q_QCASE(H)                                                                                          \
  case H:                                                                                                       \
    if (fp.method == B200Q_METHOD_QUEST) quantiser_loop<H, NV, B200Q_METHOD_QUEST, HELPER>(fp, stage, lane, stop_bar); \
    else quantiser_loop<H, NV, B200Q_METHOD_ABSMAX, HELPER>(fp, stage, lane, stop_bar);                          \
    break;
  switch (fp.had) {
    B200Q_QCASE(128)
    B200Q_QCASE(64)
    B200Q_QCASE(32)
    case 16:
      if constexpr (NV) {
        if (fp.method == B200Q_METHOD_QUEST) quantiser_loop<16, NV, B200Q_METHOD_QUEST, HELPER>(fp, stage, lane, stop_bar);
        else quantiser_loop<16, NV, B200Q_METHOD_ABSMAX, HELPER>(fp, stage, lane, stop_bar);
      }
      break;
  }
#undef B200Q_QCASE
}

// kMC (CTA pairs only): clusters of FOUR CTAs = two pairs working on N-adjacent output tiles of the same 256 rows.  The
// pairs share their A tiles: every CTA loads 64 of its 128 A rows and multicasts them to its counterpart in the other
// pair (rank ^ 2), which halves the A traffic over the L2 -> SM crossbar.  A stage may be refilled only when BOTH pairs'
// MMAs have read it (the stage-empty barriers collect one tcgen05.commit from each pair leader).
template <int kCtaGroup, int BN, bool kNV, int A_ROWS, int kF8, int kFuse, int kMC = 0>
__global__ void __launch_bounds__(kGemmThreads + 32 * kFuse, 1)
gemm_fp4_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const __grid_constant__ CUtensorMap tmap_sfa, const __grid_constant__ CUtensorMap tmap_sfb,
                const __grid_constant__ CUtensorMap tmap_d, const GemmParams p, const FuseParams fp) {
  static_assert(kFuse == 0 || ((kFuse == 2 || kFuse == 4) && kCtaGroup == 2 && A_ROWS == 128 && !kF8),
                "fused quantise+GEMM: CTA pairs, FP4 only, 2 or 4 quantiser warps");
  static_assert(kMC == 0 || (kCtaGroup == 2 && A_ROWS == 128 && kF8 != 2 && kFuse == 0), "multicast clusters: plain CTA-pair GEMM only");
  using Cfg = GemmCfg<kCtaGroup, BN, kNV, A_ROWS, kF8, kFuse>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC = Cfg::ACC_STAGES;
  constexpr int SFKB = Cfg::SFKB;
  constexpr int NB = Cfg::NB;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B swizzle atoms); identical offset in both CTAs of a pair
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stg_base + Cfg::STG_TOTAL;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + ACC + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::STG_TOTAL + 8 * (2 * STAGES + 2 * ACC));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;   // 0..1, kMC: 0..3
  const uint32_t cta_rank = crank & 1u;                                  // rank inside the CTA pair
  const uint32_t pair = kMC ? (crank >> 1) : 0u;                         // which pair of the cluster
  const uint32_t leader_rank = crank & ~1u;                              // cluster rank of this pair's leader
  const bool is_leader = cta_rank == 0;
  constexpr int kClusterCtas = kCtaGroup * (kMC ? 2 : 1);
  const int cluster_id = blockIdx.x / kClusterCtas;
  const int num_clusters = gridDim.x / kClusterCtas;
  const bool nfast = (B200Q_FLAGS(p) & 2048) != 0;   // tile walk: M-fastest (B tiles shared by concurrent clusters); profiling flag 2048: N-fastest
  // kMC: a cluster tile is two N-adjacent tiles (one per pair); an odd tile count leaves the second pair an empty tile
  const int tiles_n_eff = kMC ? (p.tiles_n + 1) / 2 : p.tiles_n;
  const int total_tiles = p.tiles_m * tiles_n_eff;
  auto tile_mn = [&](int tile, int& tm, int& tn) {
    tm = nfast ? tile / tiles_n_eff : tile % p.tiles_m;
    const int t = nfast ? tile - tm * tiles_n_eff : tile / p.tiles_m;
    tn = kMC ? t * 2 + (int)pair : t;
  };

  // a dependent grid (e.g. the tail GEMM of a split launch) may start its prologue / weight loads while this one runs
  pdl_launch_dependents();
  if (threadIdx.x == 0) { ktrace(B200Q_FLAGS(p), 0); ktrace(B200Q_FLAGS(p), 1, true); }

  // ------------------------------------------------------------------ setup
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    prefetch_tensormap(&tmap_sfa);
    prefetch_tensormap(&tmap_sfb);
    prefetch_tensormap(&tmap_d);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);     // one arrive.expect_tx by the leader's producer; bytes from both CTAs
      mbar_init(empty_bar(s), kMC ? 2 : 1);    // one tcgen05.commit (multicast to both CTAs) per pair of the cluster
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiWarps * kCtaGroup);   // every epilogue warp of the pair arrives on the leader's
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<kCtaGroup>(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish<kCtaGroup>();
  }
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
  tc_fence_after();
  // shfl from lane 0: lets the compiler treat the TMEM base as warp-uniform (uniform-datapath operands)
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_gen, 0);
  const uint32_t tmem_sfa = tmem_base + ACC * BN;
  const uint32_t tmem_sfb = tmem_sfa + Cfg::SFA_COLS;
  if (threadIdx.x == 0) ktrace(B200Q_FLAGS(p), 2);

  // ------------------------------------------------------------------ roles
  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp runs the loop with warp-uniform values; one elected lane issues.  (Issuing from a
    // divergent `lane == 0` branch makes the compiler wrap every TMA / tcgen05 op in an R2UR waterfall loop.)
    {
      const bool elected = elect_one();
      // completion goes to the leader's barrier (own barrier when kCtaGroup == 1)
      const uint32_t full0 = (kCtaGroup == 2) ? mapa(bar_base, leader_rank) : bar_base;
      const uint16_t mc_mask = (uint16_t)((1u << crank) | (1u << (crank ^ 2u)));
      int my_tiles = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) ++my_tiles;
      const int total_kt = my_tiles * p.k_tiles;
      // Programmatic dependent launch: the weights (B, SFB) do not depend on the previous kernel in the stream
      // (normally our quantise kernel producing A / SFA), so the first ring of weight loads is issued BEFORE
      // griddepcontrol.wait and overlaps the predecessor's tail; activations are loaded only after it.
      const int pre = total_kt < STAGES ? total_kt : STAGES;
      // (tile, kt) cursor advanced incrementally: no divisions on the per-k-tile path
      struct Cursor { int tile, kt, m0, n0, nb0, tm; };
      auto set_tile = [&](Cursor& c) {
        // fused: N-fastest (the first wave touches only the first row blocks of A); otherwise M-fastest
        int tm, tn;
        tile_mn(c.tile, tm, tn);
        c.tm = tm;
        c.m0 = (tm * kCtaGroup + (int)cta_rank) * BM;            // this CTA's A rows
        c.n0 = tn * BN;
        c.nb0 = c.n0 + (int)cta_rank * Cfg::B_ROWS;              // this CTA's B rows
      };
      auto advance = [&](Cursor& c) {
        if (++c.kt == p.k_tiles) { c.kt = 0; c.tile += num_clusters; set_tile(c); }
      };
      // profiling flags (timing only, wrong results): 1 << 20 skips the B tile loads, 1 << 21 the A tile loads; 1 << 22: both are
      // skipped AFTER the first ring (the stages then keep the random tiles they were filled with: the tensor pipe works on
      // realistic data with no operand traffic at all -- the power-limited FP4 ceiling of tools/fp4_peak_probe.py)
      int lflags = B200Q_FLAGS(p);
      auto load_weights = [&](int stage, int n0, int nb0, int kt) {
        const uint32_t sb = smem_base + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES;
        const uint32_t ssfb = sb + Cfg::B_BYTES + Cfg::SFA_BYTES;
        const uint32_t fb = full0 + 8u * stage;
        const uint32_t tx = Cfg::TX_BYTES - ((lflags & (1 << 20)) ? (uint32_t)Cfg::B_BYTES * kCtaGroup : 0u) -
                            ((lflags & (1 << 21)) ? (uint32_t)Cfg::A_BYTES * kCtaGroup : 0u);
        if (is_leader) mbar_arrive_expect_tx(bar_base + 8u * stage, tx);
        if (!(lflags & (1 << 20))) tma_load_2d<kCtaGroup>(sb, &tmap_b, fb, kt * BK_BYTES, nb0);
        tma_load_3d<kCtaGroup>(ssfb, &tmap_sfb, fb, 0, kt * SFKB, n0 / 128);
      };
      auto load_acts = [&](int stage, int m0, int kt) {
        const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
        const uint32_t ssfa = sa + Cfg::A_BYTES + Cfg::B_BYTES;
        const uint32_t fb = full0 + 8u * stage;
        if (lflags & (1 << 21)) {}
        else if constexpr (kMC != 0)      // my 64 rows of the A tile, to me and to my counterpart in the other pair
          tma_load_2d_multicast<2>(sa + pair * (64 * BK_BYTES), &tmap_a, fb, kt * BK_BYTES, m0 + (int)pair * 64, mc_mask);
        else if constexpr (kF8 == 2) tma_load_2d<kCtaGroup>(sa, &tmap_a, fb, m0, kt * Cfg::BK_ELEMS);   // A [K, M]: box = 128 K-rows x 128 M-bytes
        else tma_load_2d<kCtaGroup>(sa, &tmap_a, fb, kt * BK_BYTES, m0);
        tma_load_3d<kCtaGroup>(ssfa, &tmap_sfa, fb, 0, kt * SFKB, m0 / 128);
      };
      Cursor cur{cluster_id, 0, 0, 0, 0, 0};
      set_tile(cur);
      // fused: A / SFA of row block tm exist once the quantiser warps (of ALL CTAs) have written its warp-tiles
      int ready_tm = -1;
      auto wait_acts = [&](int tm) {
        if constexpr (kFuse) {
          if (tm != ready_tm && !(B200Q_FLAGS(p) & (1024 | 8192))) {     // profiling flags: 1024 no quantisers + no waits, 8192 no waits
            const int rows = (p.M - tm * 256) < 256 ? (p.M - tm * 256) : 256;
            wait_counter_ge(fp.ctr + 2 + tm, (uint32_t)rows * fp.tiles_per_row, 7);
            fence_proxy_async_global();     // generic-proxy writes (other SMs) -> this SM's TMA (async proxy) reads
            ready_tm = tm;
          }
        }
      };
      // Only with static_weights are the real loads issued early: by default B / SFB may come from the kernel right in front
      // of us (QAT re-quantises the weights every step; the reference's own pattern is quantise(a); quantise(b); matmul), so
      // every global READ waits.  What is always safe is an L2 PREFETCH of the same boxes (L2 is the coherence point: a line
      // the predecessor writes later is updated in place), so the default path still pulls its first ring of weights towards
      // the chip while the predecessor drains and then loads them from L2.
      {
        Cursor c = cur;
        for (int g = 0; g < pre; ++g) {          // ring is empty: no wait needed for the first STAGES slots
          if (elected) {
            if (p.static_weights) {
              load_weights(g, c.n0, c.nb0, c.kt);
            } else {
              if (!(lflags & (1 << 20))) tma_prefetch_2d(&tmap_b, c.kt * BK_BYTES, c.nb0);
              tma_prefetch_3d(&tmap_sfb, 0, c.kt * SFKB, c.n0 / 128);
            }
          }
          advance(c);
        }
      }
      pdl_wait();
      if (lane == 0) ktrace(B200Q_FLAGS(p), 3);
      for (int g = 0; g < pre; ++g) {
        wait_acts(cur.tm);
        if (elected) {
          if (!p.static_weights) load_weights(g, cur.n0, cur.nb0, cur.kt);
          load_acts(g, cur.m0, cur.kt);
        }
        advance(cur);
      }
      __syncwarp();
      if (B200Q_FLAGS(p) & (1 << 22)) lflags |= (3 << 20);
      int stage = (pre == STAGES) ? 0 : pre;
      uint32_t phase = (pre == STAGES) ? 1 : 0;
      for (int g = pre; g < total_kt; ++g) {
        mbar_wait(bar_base + 8u * (STAGES + stage), phase ^ 1, 1);
        if (elected) load_weights(stage, cur.n0, cur.nb0, cur.kt);
        wait_acts(cur.tm);
        if (elected) load_acts(stage, cur.m0, cur.kt);
        __syncwarp();
        advance(cur);
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (is_leader) {
      // ===================== MMA issuer =====================
      const bool elected = elect_one();   // whole warp runs the (uniform) loop; one lane issues
      // One thread issues every tcgen05 op, so its own dependent-ALU chain is on the critical path: all
      // descriptors are pre-split into a constant high word and a low word that is one add away per op.
      // (Measured: issuing the scale copies of k-tile g+1 ahead of the MMAs of k-tile g is SLOWER -- the
      // tensor pipe runs cp/mma in order -- so the order is cp(g), mma(g).)
      // instruction descriptor: e2m1 operands (format 1) for the mxf4 kinds, e4m3 (format 0) for mxf8f6f4
      constexpr uint32_t idesc_base = kF8 ? ((make_idesc_fp4(BM * kCtaGroup, BN, true) & ~((1u << 7) | (1u << 10))) |
                                             (kF8 == 2 ? (1u << 15) : 0u))                 // a_major = MN
                                          : make_idesc_fp4(BM * kCtaGroup, BN, !kNV);
      // K advance of the A descriptor per MMA: 32 bytes along the swizzled row (K-major), or 32 K-rows of 128 bytes
      // = four 1024-byte swizzle atoms (MN-major: canonical ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units, SBO = 1024 B)
      constexpr uint32_t kAStep16 = (kF8 == 2) ? (32u * 128u) >> 4 : 2u;
      constexpr uint32_t kDescHiAB = (1024u >> 4) | (1u << 14) | (kLayoutSw128 << 29);   // SBO 1024 B, version 1, 128B swizzle
      constexpr uint32_t kDescHiSF = (128u >> 4) | (1u << 14);                            // SBO 128 B, version 1, no swizzle
      constexpr uint32_t kStage16 = Cfg::STAGE_BYTES >> 4;
      const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);                  // LBO = 16 B
      const uint32_t sfa_lo0 = ((smem_base & 0x3FFFFu) >> 4) + ((Cfg::A_BYTES + Cfg::B_BYTES) >> 4);
      auto mk = [](uint32_t lo, uint32_t hi) {
        uint64_t d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
        return d;
      };
      const bool skip_cp = (B200Q_FLAGS(p) & 32) != 0;     // profiling flag 32: no scale copies (timing only)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      int tidx = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters, ++tidx) {
        int tm_, tn;
        tile_mn(tile, tm_, tn);
        const int n0 = tn * BN;
        const uint32_t sfb_shift = (uint32_t)((n0 % 128) / 32);     // 0 or 2 columns into the first SFB block
        mbar_wait(tempty_bar(acc), acc_phase ^ 1, 2);
        tc_fence_after();
        trace_event(B200Q_FLAGS(p), tidx, 0);
        if (B200Q_FLAGS(p) & (1 << 24)) {                       // profiling builds: when did this tile's first k-tile land?
          mbar_wait(bar_base + 8u * stage, phase, 9);    // (non-consuming: the issue loop waits on the same phase again)
          trace_event(B200Q_FLAGS(p), tidx, 1);
        }
        const uint32_t tmem_acc = tmem_base + acc * BN;
        const uint32_t tsfb = tmem_sfb + sfb_shift;
        // One k-tile: wait for the stage, copy its scales into TMEM (each 512-B block -> 4 columns, replicated over the 4
        // lane quarters; the copies of K-chunk c go right before the first MMA that needs them), 4 MMAs of K = 64 (32 bytes
        // = 2 x 16 B along the swizzled row each), free the stage.  This loop is the issue-rate limit of every tile narrower
        // than 256 columns (ncu / the timeline probe: ~430 cycles per k-tile vs 380 of tensor time at BN = 192), so the
        // hot instantiation (kTail = false: whole k-tile) has no K-tail predicates, no debug branches and a wait without
        // the time-out path, which lets the compiler keep the loop state in uniform registers.
        auto k_tile = [&](auto tail_tag, int kt, int k_left) {
          constexpr bool kTail = decltype(tail_tag)::value;
          mbar_wait_spin(bar_base + 8u * stage, phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + stage * kStage16;
          const uint32_t b_lo = a_lo + (Cfg::A_BYTES >> 4);
          const uint32_t sfa_lo = sfa_lo0 + stage * kStage16;
          const uint32_t sfb_lo = sfa_lo + (Cfg::SFA_BYTES >> 4);
          if (elected) {
            auto copy_chunk = [&](int b) {
              tmem_cp_32x128b_warpx4<kCtaGroup>(tmem_sfa + b * 4, mk(sfa_lo + b * 32, kDescHiSF));
#pragma unroll
              for (int nb = 0; nb < NB; ++nb)
                tmem_cp_32x128b_warpx4<kCtaGroup>(tmem_sfb + b * (4 * NB) + nb * 4,
                                                  mk(sfb_lo + (nb * SFKB + b) * 32, kDescHiSF));
            };
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              // scales per MMA: MXF4 2 bytes (sf_id 0/2 of a 4-byte cell), NVF4 a whole cell, MXF8 1 byte (sf_id = kb)
              const uint32_t chunk = kNV ? kb : (kF8 ? 0 : (kb >> 1));
              if (!skip_cp && (kNV || (kF8 ? kb == 0 : (kb & 1) == 0))) copy_chunk((int)chunk);
              if (!kTail || k_left > kb * Cfg::MMA_K) {
                const uint32_t sf_id = kNV ? 0u : (kF8 ? (uint32_t)kb : (uint32_t)((kb & 1) * 2));
                mma_fp4_block_scaled<kCtaGroup, kNV, (kF8 != 0)>(tmem_acc, mk(a_lo + kb * kAStep16, kDescHiAB), mk(b_lo + kb * 2, kDescHiAB),
                                                     idesc_base | (sf_id << 4) | (sf_id << 29), tmem_sfa + chunk * 4,
                                                     tsfb + chunk * (4 * NB), (kb > 0) ? 1u : (kt > 0 ? 1u : 0u));
              }
            }
            // stage free once these MMAs have read it (kMC: tell all four CTAs -- the other pair multicasts into our stages)
            tc_commit<kCtaGroup>(bar_base + 8u * (STAGES + stage), kMC ? (uint16_t)0xF : (uint16_t)3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        };
        const int full_kt = p.K / Cfg::BK_ELEMS;
        for (int kt = 0; kt < full_kt; ++kt) k_tile(std::false_type{}, kt, Cfg::BK_ELEMS);
        if (full_kt < p.k_tiles) k_tile(std::true_type{}, full_kt, p.K - full_kt * Cfg::BK_ELEMS);
        if (elected) tc_commit<kCtaGroup>(tfull_bar(acc), (uint16_t)(3u << leader_rank));   // accumulator complete
        __syncwarp();
        trace_event(B200Q_FLAGS(p), tidx, 2);
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (kFuse && warp >= 2 + kEpiWarps) {
    // ===================== quantiser (warps 10.., fused kernel only) =====================
    if constexpr (kFuse != 0) {
      pdl_wait();   // x comes from the previous kernel in the stream; the outputs may still be read by it
      if (!(B200Q_FLAGS(p) & 1024)) zero_fill_sf_padding(fp.q, (int64_t)blockIdx.x * (32 * kFuse) + (threadIdx.x - kGemmThreads),
                           (int64_t)gridDim.x * (32 * kFuse));
      uint4* qstage = reinterpret_cast<uint4*>(smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::STG_TOTAL + Cfg::BAR_BYTES) +
                      (warp - 2 - kEpiWarps) * 128;
      if (!(B200Q_FLAGS(p) & 1024)) quantiser_role<kNV, false>(fp, qstage, lane, 0u);   // profiling flag 1024: GEMM part only
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int ew = warp - 2;
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int half = ew >> 2;                        // which column half of the tile
    const int col0 = half * Cfg::EPI_COLS;
    const uint32_t stg = stg_base + ew * Cfg::STG_BYTES;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader = (kCtaGroup == 2) ? mapa(tempty_bar(0), leader_rank) : tempty_bar(0);
    pdl_wait();   // D must not be written before the predecessor kernel has finished (it may still read that memory)
    const float alpha = __ldg(p.alpha);
    if constexpr (kFuse != 0) {
      // Until this CTA's first accumulator is complete the 8 epilogue warps have nothing to do: they help quantise
      // (their TMA-store staging buffer doubles as the quantiser staging).  All of A is then written about as fast as
      // by the standalone kernel, and the M-fastest tile walk of the first round can start row block by row block.
      if (!(B200Q_FLAGS(p) & (1024 | 4096)))
        quantiser_role<kNV, true>(fp, reinterpret_cast<uint4*>(smem_gen + STAGES * Cfg::STAGE_BYTES + ew * Cfg::STG_BYTES), lane,
                                  tfull_bar(0));
    }
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      int tm, tn;
      tile_mn(tile, tm, tn);
      const int m0 = (tm * kCtaGroup + (int)cta_rank) * BM;
      const int n0 = tn * BN;
      mbar_wait(tfull_bar(acc), acc_phase, 6);
      tc_fence_after();
      const int tidx = (tile - cluster_id) / num_clusters;
      if (ew == 0) trace_event(B200Q_FLAGS(p), tidx, 3);
      const uint32_t taddr = tmem_base + acc * BN + col0 + ((uint32_t)(q * 32) << 16);
      if constexpr (kFuse > 2) {
        // 448 threads -> 128 registers: drain chunk by chunk, keeping only packed bf16 pairs (EPI_COLS / 2 registers)
        uint32_t pk[Cfg::EPI_COLS / 2];
#pragma unroll
        for (int ch = 0; ch < Cfg::EPI_NCHUNK; ++ch) {
          uint32_t r[Cfg::EPI_CHUNK];
#pragma unroll
          for (int j = 0; j < Cfg::EPI_CHUNK / 32; ++j) tmem_ld_32x32b_x32(taddr + ch * Cfg::EPI_CHUNK + j * 32, r + j * 32);
          tmem_ld_wait();
          if (ch == Cfg::EPI_NCHUNK - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(tempty_leader + 8u * acc);
          }
#pragma unroll
          for (int i = 0; i < Cfg::EPI_CHUNK / 2; ++i) {
            __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[2 * i]) * alpha, __uint_as_float(r[2 * i + 1]) * alpha);
            pk[ch * (Cfg::EPI_CHUNK / 2) + i] = *reinterpret_cast<uint32_t*>(&h2);
          }
        }
#pragma unroll
        for (int ch = 0; ch < Cfg::EPI_NCHUNK; ++ch) {
          if (lane == 0) bulk_wait_group_read<0>();     // staging buffer free again
          __syncwarp();
          constexpr int PIECES = Cfg::EPI_CHUNK / 8;
#pragma unroll
          for (int jj = 0; jj < PIECES; ++jj) {
            const int phys = (Cfg::EPI_CHUNK == 64) ? (jj ^ (lane & 7)) : (jj ^ ((lane >> 1) & 3));
            const uint32_t addr = stg + lane * (Cfg::EPI_CHUNK * 2) + phys * 16;
            const int b = ch * (Cfg::EPI_CHUNK / 2) + jj * 4;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[b]), "r"(pk[b + 1]), "r"(pk[b + 2]),
                         "r"(pk[b + 3])
                         : "memory");
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmap_d, stg, n0 + col0 + ch * Cfg::EPI_CHUNK, m0 + q * 32);
            bulk_commit_group();
          }
        }
        if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      // drain this warp's 32 x EPI_COLS slice of the accumulator into registers, then release TMEM at once
      uint32_t r[Cfg::EPI_COLS];
      if (!(B200Q_FLAGS(p) & 2)) {
#pragma unroll
        for (int j = 0; j < Cfg::EPI_COLS / 32; ++j) tmem_ld_32x32b_x32(taddr + j * 32, r + j * 32);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < Cfg::EPI_COLS; ++i) r[i] = 0;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCtaGroup == 2) mbar_arrive_cluster(tempty_leader + 8u * acc);
        else mbar_arrive(tempty_bar(acc));
      }
      if (ew == 0) trace_event(B200Q_FLAGS(p), tidx, 4);
      if (B200Q_FLAGS(p) & 512) __nanosleep((uint32_t)ew * (((B200Q_FLAGS(p) >> 12) & 0xff) * 50u));   // profiling: stagger the warps
      if (!(B200Q_FLAGS(p) & 1)) {
        if (p.tma_store) {
#pragma unroll
          for (int ch = 0; ch < Cfg::EPI_NCHUNK; ++ch) {
            // staging buffer must have been read by the previous TMA store
            if (lane == 0 && p.tma_store == 1) bulk_wait_group_read<0>();
            __syncwarp();
            constexpr int PIECES = Cfg::EPI_CHUNK / 8;   // 16-byte pieces per staged row (8 or 4)
#pragma unroll
            for (int jj = 0; jj < PIECES; ++jj) {
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float lo = __uint_as_float(r[ch * Cfg::EPI_CHUNK + jj * 8 + 2 * e]) * alpha;
                const float hi = __uint_as_float(r[ch * Cfg::EPI_CHUNK + jj * 8 + 2 * e + 1]) * alpha;
                __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
                w[e] = *reinterpret_cast<uint32_t*>(&h2);
              }
              // hardware swizzle of the store tensor map: 128B rows -> chunk ^= row % 8; 64B rows -> chunk ^= (row / 2) % 4
              const int phys = (Cfg::EPI_CHUNK == 64) ? (jj ^ (lane & 7)) : (jj ^ ((lane >> 1) & 3));
              const uint32_t addr = stg + lane * (Cfg::EPI_CHUNK * 2) + phys * 16;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                           : "memory");
            }
            if (p.tma_store == 2) {
              // variant: coalesced st.global from the staged tile (no TMA store): each instruction writes 4 full rows
              // segments of EPI_CHUNK*2 bytes; reads undo the staging swizzle, so they are bank-conflict free
              __syncwarp();
              constexpr int LPR = Cfg::EPI_CHUNK / 8;              // lanes (16-B pieces) per row: 8 or 4
              constexpr int RPI = 32 / LPR;                        // rows per instruction: 4 or 8
#pragma unroll
              for (int i = 0; i < 32 / RPI; ++i) {
                const int rr = i * RPI + lane / LPR, jj = lane % LPR;
                const int phys = (Cfg::EPI_CHUNK == 64) ? (jj ^ (rr & 7)) : (jj ^ ((rr >> 1) & 3));
                uint32_t w0, w1, w2, w3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                             : "r"(stg + rr * (Cfg::EPI_CHUNK * 2) + phys * 16));
                const int grow = m0 + q * 32 + rr, gcol = n0 + col0 + ch * Cfg::EPI_CHUNK + jj * 8;
                if (grow < p.M && gcol < p.N)
                  *reinterpret_cast<uint4*>(p.d + (int64_t)grow * p.ldd + gcol) = make_uint4(w0, w1, w2, w3);
              }
              __syncwarp();
              continue;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(B200Q_FLAGS(p) & 16)) {
              // profiling flag 8: every store lands on the first tile (same lines over and over: L2-resident)
              const int cx = (B200Q_FLAGS(p) & 8) ? (col0 + ch * Cfg::EPI_CHUNK) : (n0 + col0 + ch * Cfg::EPI_CHUNK);
              const int cy = (B200Q_FLAGS(p) & 8) ? (q * 32) : (m0 + q * 32);
              tma_store_2d(&tmap_d, stg, cx, cy);
              bulk_commit_group();
            }
            if (B200Q_FLAGS(p) & 256) __nanosleep(((B200Q_FLAGS(p) >> 12) & 0xff) * 50u);   // profiling: pace the stores
          }
        } else if (p.tma_store == 0 && (p.ldd % 16) == 0) {
          // direct 256-bit stores: thread = row, one full 32-byte sector per instruction
          const int row = m0 + q * 32 + lane;
          if (row < p.M) {
            __nv_bfloat16* drow = p.d + (int64_t)row * p.ldd + n0 + col0;
#pragma unroll
            for (int v = 0; v < Cfg::EPI_COLS / 16; ++v) {
              if (n0 + col0 + v * 16 < p.N) {
                uint32_t w[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                  __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(r[v * 16 + 2 * e]) * alpha,
                                                            __uint_as_float(r[v * 16 + 2 * e + 1]) * alpha);
                  w[e] = *reinterpret_cast<uint32_t*>(&h2);
                }
                asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(drow + v * 16), "r"(w[0]),
                             "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                             : "memory");
              }
            }
          }
        } else {
          // scalar stores (row pitch not a multiple of 16 bytes, i.e. N % 8 != 0): thread = row
          const int row = m0 + q * 32 + lane;
          if (row < p.M) {
            uint16_t* drow = reinterpret_cast<uint16_t*>(p.d) + (int64_t)row * p.ldd + n0 + col0;
#pragma unroll
            for (int i = 0; i < Cfg::EPI_COLS; ++i) {
              if (n0 + col0 + i < p.N) {
                __nv_bfloat16 h = __float2bfloat16_rn(__uint_as_float(r[i]) * alpha);
                drow[i] = *reinterpret_cast<uint16_t*>(&h);
              }
            }
          }
        }
      }
      if (ew == 0) trace_event(B200Q_FLAGS(p), tidx, 5);
      if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait_group<0>();   // all TMA stores of this warp have completed
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
  if (threadIdx.x == 0) { ktrace(B200Q_FLAGS(p), 5); ktrace(B200Q_FLAGS(p), 6, true); }
  if (warp == 1) {
    __syncwarp();   // .sync.aligned: the issuing lane must have reconverged with its warp
    tmem_dealloc<kCtaGroup>(tmem_base, Cfg::TMEM_COLS);
  }
  if constexpr (kFuse != 0) {
    // every poll / release of this CTA is behind the barrier above; the last CTA to get here re-zeroes the workspace
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(fp.ctr + 1, 1u) == gridDim.x - 1) {
        for (int i = 0; i < p.tiles_m; ++i) fp.ctr[2 + i] = 0u;
        fp.ctr[0] = 0u;
        fp.ctr[1] = 0u;
        __threadfence();
      }
    }
  }
}

// ------------------------------------------------------------------ hybrid tile pairs (PROFILING BUILDS ONLY, configuration (2, 448))
// Measured on B200 at the start of round 2 (profiles/r02_notes.md): bit-identical, and NOT faster -- config 1 GEMM 5380 vs
// 5449 TFLOP/s (MX), 5066 vs 5076 (NV) -- although it removes the accumulator hand-off bubble of the (2,256) tile completely.
// The part is power-limited under FP4 MMA load: closing an idle gap lowers the clock instead of the run time.  The kernel is
// therefore NOT part of the product library; it stays in the profiling build (-DB200Q_PROFILING) as the bubble-free control
// of the tensor-rate measurements (tools/fp4_peak_probe.py).
#ifdef B200Q_PROFILING
// Every CTA pair walks "super tiles" of 256 rows x 448 columns as ONE 256-wide and ONE 192-wide tile.  A 256-column and a
// 192-column fp32 accumulator plus the scale columns fit the 512 TMEM columns (256 + 192 + 24 / 48), so -- unlike the plain
// (2,256) configuration, whose single accumulator costs ~1800 of ~10000 cycles per tile at the hand-off --
// the MMA warp always finds a free accumulator: the tensor pipe never waits for a drain.  14336 = 32 x 448 and 28672 = 64 x 448,
// so the Llama FFN shapes also lose the 11 %-full last round of tiles (512 super tiles over 74 pairs = 6.92 rounds).
// Column order inside two neighbouring super tiles is W N | N W: every wide tile then starts on a 128-row scale block and
// every narrow tile 0 or 64 rows into one (an even TMEM column shift; odd shifts fault, profiles/r01_notes.md).
// Requires N % 448 == 0, K % 256 == 0, ldd % 8 == 0, FP4 kinds.  Same arithmetic as every other configuration (one fp32 chain
// per output over K, alpha once, one RNE).  STATUS: written after round 1's GPU budget was spent -- compiled, NOT yet run;
// reachable only through the explicit configuration or B200Q_GEMM_HYBRID=1.
// super tile `sup` (row blocks fastest), half j (0 = the 256-wide tile, 1 = the 192-wide one) -> row block tm, first column n0.
// Shared by the kernel and the host-only query b200q_debug_hybrid_tile (tests/test_cabi.py checks coverage and alignment).
__host__ __device__ __forceinline__ void hybrid_tile_geom(int sup, int j, int tiles_m, int& tm, int& n0) {
  tm = sup % tiles_m;
  const int sn = sup / tiles_m;
  const int base = sn * 448;
  n0 = (sn & 1) ? (j == 0 ? base + 192 : base) : (j == 0 ? base : base + 256);
}

template <bool kNV>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_fp4_hybrid_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_bw,
                       const __grid_constant__ CUtensorMap tmap_bn, const __grid_constant__ CUtensorMap tmap_sfa,
                       const __grid_constant__ CUtensorMap tmap_sfb, const __grid_constant__ CUtensorMap tmap_d64,
                       const __grid_constant__ CUtensorMap tmap_d32, const GemmParams p) {
  using Cfg = GemmCfg<2, 256, kNV>;          // stage layout of the (2,256) configuration: the narrow tile uses 96 of the 128 B rows
  constexpr int STAGES = Cfg::STAGES;
  constexpr int SFKB = Cfg::SFKB;
  constexpr int NB = 2;                      // SFB row-blocks a tile touches (wide: 2 whole blocks; narrow: 192 rows from row 0 or 64)
  constexpr int BNW = 256, BNN = 192, SUPER = BNW + BNN;
  constexpr int ACC = 2;
  static_assert(Cfg::NB == NB, "scale staging of the (2,256) stage");
  static_assert(SUPER + Cfg::SF_COLS <= 512, "TMEM: wide + narrow accumulator + scales");

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t stg_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = stg_base + Cfg::STG_TOTAL;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + ACC + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 2 * ACC);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(
      smem_gen + STAGES * Cfg::STAGE_BYTES + Cfg::STG_TOTAL + 8 * (2 * STAGES + 2 * ACC));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();          // 0 = leader of the pair
  const bool is_leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int total_supers = p.tiles_m * p.tiles_n;        // tiles_n = N / 448 here
  // super tile `sup` (M fastest), half j (0 = wide, 1 = narrow) -> row block tm, first column n0
  static_assert(SUPER == 448 && BNN == 192 && BNW == 256, "hybrid_tile_geom");
  auto geom = [&](int sup, int j, int& tm, int& n0) { hybrid_tile_geom(sup, j, p.tiles_m, tm, n0); };

  pdl_launch_dependents();

  // ------------------------------------------------------------------ setup
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_bw);
    prefetch_tensormap(&tmap_bn);
    prefetch_tensormap(&tmap_sfa);
    prefetch_tensormap(&tmap_sfb);
    prefetch_tensormap(&tmap_d64);
    prefetch_tensormap(&tmap_d32);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), kEpiWarps * 2);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<2>(tmem_slot, 512);
    tmem_relinquish<2>();
  }
  tc_fence_before();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_gen, 0);
  const uint32_t tmem_sfa = tmem_base + SUPER;
  const uint32_t tmem_sfb = tmem_sfa + Cfg::SFA_COLS;

  if (warp == 0) {
    // ===================== TMA producer (same protocol as gemm_fp4_kernel) =====================
    const bool elected = elect_one();
    const uint32_t full0 = mapa(bar_base, 0);
    int my_supers = 0;
    for (int sup = cluster_id; sup < total_supers; sup += num_clusters) ++my_supers;
    const int total_kt = my_supers * 2 * p.k_tiles;
    const int pre = total_kt < STAGES ? total_kt : STAGES;
    struct Cursor { int sup, j, kt, m0, n0, nb0; };
    auto set_tile = [&](Cursor& c) {
      int tm, n0;
      geom(c.sup, c.j, tm, n0);
      c.m0 = (tm * 2 + (int)cta_rank) * BM;
      c.n0 = n0;
      c.nb0 = n0 + (int)cta_rank * ((c.j ? BNN : BNW) / 2);
    };
    auto advance = [&](Cursor& c) {
      if (++c.kt == p.k_tiles) {
        c.kt = 0;
        if (c.j == 0) { c.j = 1; } else { c.j = 0; c.sup += num_clusters; }
        set_tile(c);
      }
    };
    auto load_weights = [&](int stage, const Cursor& c) {
      const uint32_t sb = smem_base + stage * Cfg::STAGE_BYTES + Cfg::A_BYTES;
      const uint32_t ssfb = sb + Cfg::B_BYTES + Cfg::SFA_BYTES;
      const uint32_t fb = full0 + 8u * stage;
      const uint32_t b_bytes = (uint32_t)((c.j ? BNN : BNW) / 2) * BK_BYTES;
      const uint32_t tx = 2u * ((uint32_t)Cfg::A_BYTES + b_bytes + (uint32_t)Cfg::SFA_BYTES + (uint32_t)Cfg::SFB_BYTES);
      if (is_leader) mbar_arrive_expect_tx(bar_base + 8u * stage, tx);
      if (c.j) tma_load_2d<2>(sb, &tmap_bn, fb, c.kt * BK_BYTES, c.nb0);
      else tma_load_2d<2>(sb, &tmap_bw, fb, c.kt * BK_BYTES, c.nb0);
      tma_load_3d<2>(ssfb, &tmap_sfb, fb, 0, c.kt * SFKB, c.n0 / 128);
    };
    auto load_acts = [&](int stage, const Cursor& c) {
      const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
      const uint32_t ssfa = sa + Cfg::A_BYTES + Cfg::B_BYTES;
      const uint32_t fb = full0 + 8u * stage;
      tma_load_2d<2>(sa, &tmap_a, fb, c.kt * BK_BYTES, c.m0);
      tma_load_3d<2>(ssfa, &tmap_sfa, fb, 0, c.kt * SFKB, c.m0 / 128);
    };
    Cursor cur{cluster_id, 0, 0, 0, 0, 0};
    set_tile(cur);
    if (!p.static_weights) pdl_wait();         // see gemm_fp4_kernel: weights are prefetched early only on the caller's word
    {
      Cursor c = cur;
      for (int g = 0; g < pre; ++g) {
        if (elected) load_weights(g, c);
        advance(c);
      }
    }
    pdl_wait();
    for (int g = 0; g < pre; ++g) {
      if (elected) load_acts(g, cur);
      advance(cur);
    }
    __syncwarp();
    int stage = (pre == STAGES) ? 0 : pre;
    uint32_t phase = (pre == STAGES) ? 1 : 0;
    for (int g = pre; g < total_kt; ++g) {
      mbar_wait(bar_base + 8u * (STAGES + stage), phase ^ 1, 1);
      if (elected) {
        load_weights(stage, cur);
        load_acts(stage, cur);
      }
      __syncwarp();
      advance(cur);
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    if (is_leader) {
      // ===================== MMA issuer =====================
      const bool elected = elect_one();
      constexpr uint32_t idesc_w = make_idesc_fp4(BM * 2, BNW, !kNV);
      constexpr uint32_t idesc_n = make_idesc_fp4(BM * 2, BNN, !kNV);
      constexpr uint32_t kDescHiAB = (1024u >> 4) | (1u << 14) | (kLayoutSw128 << 29);
      constexpr uint32_t kDescHiSF = (128u >> 4) | (1u << 14);
      constexpr uint32_t kStage16 = Cfg::STAGE_BYTES >> 4;
      const uint32_t a_lo0 = ((smem_base & 0x3FFFFu) >> 4) | (1u << 16);
      const uint32_t sfa_lo0 = ((smem_base & 0x3FFFFu) >> 4) + ((Cfg::A_BYTES + Cfg::B_BYTES) >> 4);
      auto mk = [](uint32_t lo, uint32_t hi) {
        uint64_t d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
        return d;
      };
      int stage = 0;
      uint32_t phase = 0;
      uint32_t acc_phase0 = 0, acc_phase1 = 0;
      for (int sup = cluster_id; sup < total_supers; sup += num_clusters) {
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          int tm_, n0;
          geom(sup, j, tm_, n0);
          const uint32_t sfb_shift = (uint32_t)((n0 % 128) / 32);     // 0 or 2
          mbar_wait(tempty_bar(j), (j ? acc_phase1 : acc_phase0) ^ 1, 2);
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + (j ? (uint32_t)BNW : 0u);
          const uint32_t tsfb = tmem_sfb + sfb_shift;
          const uint32_t idesc_base = j ? idesc_n : idesc_w;
          for (int kt = 0; kt < p.k_tiles; ++kt) {
            mbar_wait(bar_base + 8u * stage, phase, 3);   // time-out variant until the kernel has been proven on hardware (a lost arrival traps instead of hanging)
            tc_fence_after();
            const uint32_t a_lo = a_lo0 + stage * kStage16;
            const uint32_t b_lo = a_lo + (Cfg::A_BYTES >> 4);
            const uint32_t sfa_lo = sfa_lo0 + stage * kStage16;
            const uint32_t sfb_lo = sfa_lo + (Cfg::SFA_BYTES >> 4);
            if (elected) {
              auto copy_chunk = [&](int b) {
                tmem_cp_32x128b_warpx4<2>(tmem_sfa + b * 4, mk(sfa_lo + b * 32, kDescHiSF));
#pragma unroll
                for (int nb = 0; nb < NB; ++nb)
                  tmem_cp_32x128b_warpx4<2>(tmem_sfb + b * (4 * NB) + nb * 4, mk(sfb_lo + (nb * SFKB + b) * 32, kDescHiSF));
              };
#pragma unroll
              for (int kb = 0; kb < 4; ++kb) {
                const uint32_t chunk = kNV ? kb : (kb >> 1);
                if (kNV || (kb & 1) == 0) copy_chunk((int)chunk);
                const uint32_t sf_id = kNV ? 0u : (uint32_t)((kb & 1) * 2);
                mma_fp4_block_scaled<2, kNV, false>(tmem_acc, mk(a_lo + kb * 2, kDescHiAB), mk(b_lo + kb * 2, kDescHiAB),
                                                    idesc_base | (sf_id << 4) | (sf_id << 29), tmem_sfa + chunk * 4,
                                                    tsfb + chunk * (4 * NB), (kb > 0) ? 1u : (kt > 0 ? 1u : 0u));
              }
              tc_commit<2>(bar_base + 8u * (STAGES + stage), (uint16_t)3);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (elected) tc_commit<2>(tfull_bar(j), (uint16_t)3);
          __syncwarp();
          if (j) acc_phase1 ^= 1; else acc_phase0 ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9): 4 TMEM lane quarters x 2 column halves of the tile =====================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int half = ew >> 2;
    const uint32_t stg = stg_base + ew * Cfg::STG_BYTES;
    uint32_t acc_phase0 = 0, acc_phase1 = 0;
    const uint32_t tempty_leader = mapa(tempty_bar(0), 0);
    pdl_wait();
    const float alpha = __ldg(p.alpha);
    // convert + stage + TMA-store one chunk of CH (64 or 32) columns held in rr[0 .. CH)
    auto store_chunk = [&](auto ch_tag, const uint32_t* rr, int cx, int cy) {
      constexpr int CH = decltype(ch_tag)::value;
      constexpr int PIECES = CH / 8;
      if (lane == 0) bulk_wait_group_read<0>();      // staging buffer free again
      __syncwarp();
#pragma unroll
      for (int jj = 0; jj < PIECES; ++jj) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          __nv_bfloat162 h2 = __floats2bfloat162_rn(__uint_as_float(rr[jj * 8 + 2 * e]) * alpha,
                                                    __uint_as_float(rr[jj * 8 + 2 * e + 1]) * alpha);
          w[e] = *reinterpret_cast<uint32_t*>(&h2);
        }
        const int phys = (CH == 64) ? (jj ^ (lane & 7)) : (jj ^ ((lane >> 1) & 3));
        const uint32_t addr = stg + lane * (CH * 2) + phys * 16;
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        if constexpr (CH == 64) tma_store_2d(&tmap_d64, stg, cx, cy);
        else tma_store_2d(&tmap_d32, stg, cx, cy);
        bulk_commit_group();
      }
    };
    for (int sup = cluster_id; sup < total_supers; sup += num_clusters) {
#pragma unroll 1
      for (int j = 0; j < 2; ++j) {
        int tm, n0;
        geom(sup, j, tm, n0);
        const int m0 = (tm * 2 + (int)cta_rank) * BM;
        const int cols = (j ? BNN : BNW) / 2;            // 128 or 96 columns per warp
        const int col0 = half * cols;
        mbar_wait(tfull_bar(j), j ? acc_phase1 : acc_phase0, 6);
        tc_fence_after();
        const uint32_t taddr = tmem_base + (j ? (uint32_t)BNW : 0u) + (uint32_t)col0 + ((uint32_t)(q * 32) << 16);
        uint32_t r[128];
        tmem_ld_32x32b_x32(taddr, r);
        tmem_ld_32x32b_x32(taddr + 32, r + 32);
        tmem_ld_32x32b_x32(taddr + 64, r + 64);
        if (j == 0) tmem_ld_32x32b_x32(taddr + 96, r + 96);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(tempty_leader + 8u * j);      // accumulator j is free again
        if (j) acc_phase1 ^= 1; else acc_phase0 ^= 1;
        store_chunk(std::integral_constant<int, 64>{}, r, n0 + col0, m0 + q * 32);
        if (j == 0) store_chunk(std::integral_constant<int, 64>{}, r + 64, n0 + col0 + 64, m0 + q * 32);
        else store_chunk(std::integral_constant<int, 32>{}, r + 64, n0 + col0 + 64, m0 + q * 32);
      }
    }
    if (lane == 0) bulk_wait_group<0>();
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  cluster_sync();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<2>(tmem_base, 512);
  }
}
#endif  // B200Q_PROFILING

// ------------------------------------------------------------------ host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

// A tensor map is a pure function of (address, type, dims, strides, box, swizzle): encodings are memoised per thread in a
// small 4-way set-associative table keyed on exactly those values, so a steady-state caller (same buffers every step, the
// reference benchmarks' pattern) pays five table look-ups per GEMM instead of five driver calls.  Never stale: the key
// IS the content.  (Direct-mapped until session 3: two of the five maps of ONE call could share a slot and evict each other
// on every call -- observed as 3 hits / 2 misses in test_tensor_map_cache_hits_on_repeated_calls...)  B200Q_NO_TMAP_CACHE=1
// bypasses it.
struct TmapKey {
  const void* ptr;
  uint64_t dims[3], strides[2];
  uint32_t box[3];
  uint32_t dt, rank, sw;
  bool operator==(const TmapKey& o) const {
    return ptr == o.ptr && dt == o.dt && rank == o.rank && sw == o.sw && dims[0] == o.dims[0] && dims[1] == o.dims[1] &&
           dims[2] == o.dims[2] && strides[0] == o.strides[0] && strides[1] == o.strides[1] && box[0] == o.box[0] &&
           box[1] == o.box[1] && box[2] == o.box[2];
  }
};
struct alignas(64) TmapEntry {
  CUtensorMap map;
  TmapKey key;
  bool valid;
};
constexpr int kTmapCacheSets = 64, kTmapCacheWays = 4;
static thread_local TmapEntry g_tmap_cache[kTmapCacheSets][kTmapCacheWays];
static thread_local unsigned char g_tmap_victim[kTmapCacheSets];
static thread_local unsigned long long g_tmap_hits = 0, g_tmap_misses = 0;

static int encode(CUtensorMap* tm, CUtensorMapDataType dt, int rank, const void* ptr, const cuuint64_t* dims,
                  const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle sw, const char* what) {
  TmapKey key = {};
  key.ptr = ptr;
  key.dt = (uint32_t)dt; key.rank = (uint32_t)rank; key.sw = (uint32_t)sw;
  for (int i = 0; i < rank; ++i) { key.dims[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.strides[i] = strides[i];
  uint64_t h = (uint64_t)(uintptr_t)ptr * 0x9E3779B97F4A7C15ull;
  h ^= (key.dims[0] * 31 + key.dims[1]) * 0xC2B2AE3D27D4EB4Full + key.box[1] * 0x165667B19E3779F9ull + key.box[0] + key.dims[2] * 7 + key.dt * 131 + key.sw;
  const unsigned set = (unsigned)((h >> 32) % kTmapCacheSets);
  TmapEntry* ways = g_tmap_cache[set];
  const bool use_cache = !env().no_tmap_cache;
  if (use_cache) {
    for (int w = 0; w < kTmapCacheWays; ++w) {
      if (ways[w].valid && ways[w].key == key) {
        *tm = ways[w].map;
        ++g_tmap_hits;
        return 0;
      }
    }
  }
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return B200Q_ECUDA;
  }
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, dt, (cuuint32_t)rank, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
    return B200Q_ECUDA;
  }
  ++g_tmap_misses;
  if (use_cache) {
    int w = 0;
    while (w < kTmapCacheWays && ways[w].valid) ++w;                 // a free way first, else round-robin within the set
    if (w == kTmapCacheWays) w = (g_tmap_victim[set]++) % kTmapCacheWays;
    ways[w].map = *tm;
    ways[w].key = key;
    ways[w].valid = true;
  }
  return 0;
}

// bf16 [rows, 128] view of a flat activation tensor, box = [128 rows, 64 elements], 128B swizzle (quantize_tc.cu)
int make_x128_tmap(void* tm, const void* ptr, int64_t rows) {
  cuuint64_t dims[2] = {128, (cuuint64_t)rows};
  cuuint64_t strides[1] = {256};
  cuuint32_t box[2] = {64, 128};
  return encode((CUtensorMap*)tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "x[rows,128]");
}

// bf16 [B, N, M] (M contiguous) as the 3-D tensor {M, N, B}, box = {64 columns, 128 rows, 1}, 128B swizzle: a tile in INPUT
// orientation for the transposing tensor-core quantiser (backward_tc.cu); rows >= N / columns >= M of a batch read as zero
int make_xT_tmap(void* tm, const void* ptr, int64_t M, int64_t N, int64_t B) {
  cuuint64_t dims[3] = {(cuuint64_t)M, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)M * 2, (cuuint64_t)M * (cuuint64_t)N * 2};
  cuuint32_t box[3] = {64, 128, 1};
  return encode((CUtensorMap*)tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, "x[B,N,M]");
}

// bf16 rotation matrix [H, H] (row k, column n contiguous), box = [min(H, 64) columns, H rows], swizzle span = box row
int make_rot_tmap(void* tm, const void* ptr, int had) {
  const int bw = had < 64 ? had : 64;
  cuuint64_t dims[2] = {(cuuint64_t)had, (cuuint64_t)had};
  cuuint64_t strides[1] = {(cuuint64_t)had * 2};
  cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)had};
  const CUtensorMapSwizzle sw = bw == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bw == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  return encode((CUtensorMap*)tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, sw, "rotation[H,H]");
}

// [rows, row_bytes] uint8 operand, box = [box_rows, 128 bytes], 128B swizzle, zero fill out of bounds
int make_operand_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t row_bytes, int box_rows,
                      const char* what) {
  cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)BK_BYTES, (cuuint32_t)box_rows};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, what);
}

// [rows, row_bytes] uint8 operand viewed as {128 bytes, rows, k-tiles}: ONE box = {128 B, box_rows, all k-tiles} lands in shared
// memory as consecutive 128B-swizzled [box_rows x 128 B] slabs, one per k-tile (gemm_decode.cu: the resident activations)
int make_operand_ktile_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t row_bytes, int box_rows, const char* what) {
  cuuint64_t dims[3] = {(cuuint64_t)BK_BYTES, (cuuint64_t)rows, (cuuint64_t)(row_bytes / BK_BYTES)};
  cuuint64_t strides[2] = {(cuuint64_t)row_bytes, (cuuint64_t)BK_BYTES};
  cuuint32_t box[3] = {(cuuint32_t)BK_BYTES, (cuuint32_t)box_rows, (cuuint32_t)(row_bytes / BK_BYTES)};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, what);
}

// blocked scale buffer as a 3-D tensor {128 x u32 (one 512-B block), col_blocks, row_blocks};
// box = {128, kblocks, rblocks}; out-of-range blocks read as zero (scale 2^-127 / 0.0: never NaN)
int make_sf_tmap(CUtensorMap* tm, const void* ptr, int64_t row_blocks, int64_t col_blocks, int box_kb,
                 int box_rb, const char* what) {
  cuuint64_t dims[3] = {128, (cuuint64_t)col_blocks, (cuuint64_t)row_blocks};
  cuuint64_t strides[2] = {512, (cuuint64_t)col_blocks * 512};
  cuuint32_t box[3] = {128, (cuuint32_t)box_kb, (cuuint32_t)box_rb};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE, what);
}

// D [M, N] bf16 row-major, box = [32 rows, chunk cols], swizzle matching the staging layout
static int make_d_tmap(CUtensorMap* tm, const void* ptr, int64_t M, int64_t N, int64_t ldd, int chunk) {
  cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
  cuuint64_t strides[1] = {(cuuint64_t)ldd * 2};
  cuuint32_t box[2] = {(cuuint32_t)chunk, 32};
  return encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box,
                chunk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, "D");
}

template <int kCtaGroup, int BN, bool kNV, int A_ROWS = 128, int kF8 = 0, int kFuse = 0, int kMC = 0>
static int launch_gemm(const void* A, const void* B, const void* SFA, const void* SFB, const float* alpha, void* D,
                       int M, int N, int K, int ldd, cudaStream_t stream, bool static_w, const FuseParams* fuse = nullptr) {
  using Cfg = GemmCfg<kCtaGroup, BN, kNV, A_ROWS, kF8, kFuse>;
  auto kern = gemm_fp4_kernel<kCtaGroup, BN, kNV, A_ROWS, kF8, kFuse, kMC>;
  constexpr int kClusterCtas = kCtaGroup * (kMC ? 2 : 1);
  static std::atomic<unsigned long long> smem_attr_done{0};   // per instantiation, one bit per device
  if (int rc_attr = ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, smem_attr_done)) return rc_attr;
  const int group = kNV ? 16 : 32;
  const int64_t sf_col_blocks = ceil_div(ceil_div(K, group), 4);
  CUtensorMap ta, tb, tsa, tsb, td;
  int rc;
  const int64_t row_bytes = kF8 ? K : K / 2;
  if constexpr (kF8 == 2) {
    if ((rc = make_operand_tmap(&ta, A, K, M, 128, "A[K,M]"))) return rc;      // rows = K, row pitch = M bytes, box 128 x 128
  } else {
    if ((rc = make_operand_tmap(&ta, A, M, row_bytes, kMC ? 64 : A_ROWS, "A"))) return rc;   // kMC: half tiles, multicast
  }
  if ((rc = make_operand_tmap(&tb, B, N, row_bytes, Cfg::B_ROWS, "B"))) return rc;
  if ((rc = make_sf_tmap(&tsa, SFA, ceil_div(M, 128), sf_col_blocks, Cfg::SFKB, 1, "SFA"))) return rc;
  if ((rc = make_sf_tmap(&tsb, SFB, ceil_div(N, 128), sf_col_blocks, Cfg::SFKB, Cfg::NB, "SFB"))) return rc;
  GemmParams p;
  p.alpha = alpha;
  p.d = (__nv_bfloat16*)D;
  p.M = M; p.N = N; p.K = K;
  p.ldd = ldd;
  p.tiles_m = (int)ceil_div(M, BM * kCtaGroup);
  p.tiles_n = (int)ceil_div(N, BN);
  p.k_tiles = (int)ceil_div(K, Cfg::BK_ELEMS);
  p.tma_store = (ldd % 8 == 0) ? 1 : 0;
  p.static_weights = static_w ? 1 : 0;
  p.flags = env().gemm_flags;                              // always 0 unless the library was built with -DB200Q_PROFILING
  if (B200Q_FLAGS(p) & 4) p.tma_store = 0;
  if ((B200Q_FLAGS(p) & 128) && p.tma_store) p.tma_store = 2;     // coalesced st.global from the staged tile
  if (p.tma_store) {
    if ((rc = make_d_tmap(&td, D, M, N, ldd, Cfg::EPI_CHUNK))) return rc;
  } else {
    td = ta;  // unused
  }
  const int total = p.tiles_m * (kMC ? (p.tiles_n + 1) / 2 : p.tiles_n);
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(Cfg::THREADS);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kClusterCtas;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  int clusters = num_sms() / kClusterCtas;
  if constexpr (kMC != 0) {
    // clusters of 4 must sit inside one GPC: ask how many are co-resident (a persistent grid must not exceed that)
    static std::atomic<int> max_clusters[64];     // per device (zero-initialised); racing first calls store the same value
    std::atomic<int>& mc_a = max_clusters[current_device() & 63];
    int mc = mc_a.load(std::memory_order_acquire);
    if (mc == 0) {
      cfg.gridDim = dim3((unsigned)(clusters * kClusterCtas));
      cfg.attrs = attrs;
      cfg.numAttrs = 1;
      int n = 0;
      B200Q_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      mc = n > 0 ? n : 1;
      mc_a.store(mc, std::memory_order_release);
      if (env().verbose) fprintf(stderr, "b200q: clusters of %d CTAs co-resident: %d (SMs %d)\n", kClusterCtas, n, num_sms());
    }
    if (clusters > mc) clusters = mc;
  }
  if (clusters > total) clusters = total;
  cfg.gridDim = dim3((unsigned)(clusters * kClusterCtas));
  // programmatic dependent launch: this grid may start while the previous kernel in the stream drains; everything that
  // depends on that kernel sits behind griddepcontrol.wait inside (B200Q_NO_PDL=1 disables the attribute)
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = env().no_pdl == 1 ? 1 : 2;
  FuseParams fpv = {};
  if (fuse) fpv = *fuse;
  B200Q_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, tsa, tsb, td, p, fpv));
  return 0;
}

#ifdef B200Q_PROFILING
// hybrid tile pairs (gemm_fp4_hybrid_kernel): explicit configuration (2, 448)
static bool hybrid_eligible(int M, int N, int K, int ldd, int kind) {
  return (kind == B200Q_KIND_MXF4 || kind == B200Q_KIND_NVF4) && M > 0 && N % 448 == 0 && K % 256 == 0 && ldd % 8 == 0;
}

template <bool kNV>
static int launch_gemm_hybrid(const void* A, const void* B, const void* SFA, const void* SFB, const float* alpha, void* D,
                              int M, int N, int K, int ldd, cudaStream_t stream, bool static_w) {
  using Cfg = GemmCfg<2, 256, kNV>;
  auto kern = gemm_fp4_hybrid_kernel<kNV>;
  static std::atomic<unsigned long long> smem_attr_done{0};   // per instantiation, one bit per device
  if (int rc_attr = ensure_dynamic_smem(kern, Cfg::SMEM_BYTES, smem_attr_done)) return rc_attr;
  const int group = kNV ? 16 : 32;
  const int64_t sf_col_blocks = ceil_div(ceil_div(K, group), 4);
  CUtensorMap ta, tbw, tbn, tsa, tsb, td64, td32;
  int rc;
  if ((rc = make_operand_tmap(&ta, A, M, K / 2, 128, "A"))) return rc;
  if ((rc = make_operand_tmap(&tbw, B, N, K / 2, 128, "B (wide tile)"))) return rc;
  if ((rc = make_operand_tmap(&tbn, B, N, K / 2, 96, "B (narrow tile)"))) return rc;
  if ((rc = make_sf_tmap(&tsa, SFA, ceil_div(M, 128), sf_col_blocks, Cfg::SFKB, 1, "SFA"))) return rc;
  if ((rc = make_sf_tmap(&tsb, SFB, ceil_div(N, 128), sf_col_blocks, Cfg::SFKB, 2, "SFB"))) return rc;
  if ((rc = make_d_tmap(&td64, D, M, N, ldd, 64))) return rc;
  if ((rc = make_d_tmap(&td32, D, M, N, ldd, 32))) return rc;
  GemmParams p;
  p.alpha = alpha;
  p.d = (__nv_bfloat16*)D;
  p.M = M; p.N = N; p.K = K;
  p.ldd = ldd;
  p.tiles_m = (int)ceil_div(M, BM * 2);
  p.tiles_n = N / 448;                    // super tiles along N
  p.k_tiles = K / 256;
  p.tma_store = 1;
  p.flags = 0;
  p.static_weights = static_w ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = 2;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  int clusters = num_sms() / 2;
  const int total = p.tiles_m * p.tiles_n;
  if (clusters > total) clusters = total;
  cfg.gridDim = dim3((unsigned)(clusters * 2));
  attrs[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attrs[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = env().no_pdl == 1 ? 1 : 2;
  B200Q_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tbw, tbn, tsa, tsb, td64, td32, p));
  return 0;
}
#endif  // B200Q_PROFILING

template <bool kNV, int kF8>
static int dispatch_cfg(int cta_group, int block_n, const void* A, const void* B, const void* SFA, const void* SFB,
                        const float* alpha, void* D, int M, int N, int K, int ldd, cudaStream_t s, bool sw) {
  // decode (M <= 32): the weight-streaming kernel with swapped operands (gemm_decode.cu), configuration (1, 16)
  if constexpr (kF8 == 0) {
    if (cta_group == 1 && block_n == 16)
      return launch_gemm_decode(A, B, SFA, SFB, alpha, D, M, N, K, ldd, kNV ? B200Q_KIND_NVF4 : B200Q_KIND_MXF4, sw, s);
  }
  // small M: same 128-wide single-CTA tile, fewer A rows staged (more weight k-tiles in flight)
  if constexpr (kF8 != 2) {
    if (cta_group == 1 && block_n == 128 && M <= 16) return launch_gemm<1, 128, kNV, 16, kF8>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, s, sw);
    if (cta_group == 1 && block_n == 128 && M <= 32) return launch_gemm<1, 128, kNV, 32, kF8>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, s, sw);
    if (cta_group == 1 && block_n == 128 && M <= 64) return launch_gemm<1, 128, kNV, 64, kF8>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, s, sw);
  }
#ifdef B200Q_PROFILING
  if constexpr (kF8 == 0) {
    if (cta_group == 2 && block_n == 448) {      // hybrid 256 + 192 tile pairs (profiling builds only: measured not faster)
      if (!hybrid_eligible(M, N, K, ldd, kNV ? B200Q_KIND_NVF4 : B200Q_KIND_MXF4)) {
        set_error("configuration (2, 448) needs N %% 448 == 0, K %% 256 == 0 and a row pitch that is a multiple of 8 (N=%d K=%d ldd=%d)", N, K, ldd);
        return B200Q_EINVAL;
      }
      return launch_gemm_hybrid<kNV>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, s, sw);
    }
  }
#endif
#define B200Q_CASE(CG, BNV) \
  if (cta_group == CG && block_n == BNV) return launch_gemm<CG, BNV, kNV, 128, kF8>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, s, sw);
  B200Q_CASE(1, 128)
  B200Q_CASE(1, 256)
  B200Q_CASE(2, 128)
  B200Q_CASE(2, 192)
  B200Q_CASE(2, 256)
  if constexpr (kF8 == 0) {
    // cta_group 4 = CTA pairs in clusters of four, A tiles multicast between the two pairs
    if (cta_group == 4 && block_n == 256) return launch_gemm<2, 256, kNV, 128, 0, 0, 1>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, s, sw);
    if (cta_group == 4 && block_n == 192) return launch_gemm<2, 192, kNV, 128, 0, 0, 1>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, s, sw);
  }
  if constexpr (kF8 == 0) {
    B200Q_CASE(1, 64)
    B200Q_CASE(1, 192)
  }
#undef B200Q_CASE
  set_error("unsupported GEMM configuration cta_group=%d block_n=%d", cta_group, block_n);
  return B200Q_EINVAL;
}

}  // namespace b200q

namespace b200q {
struct GemmPlan { int cta_group, block_n, n_main; };   // n_main > 0: peel columns [n_main, N) into a second launch
static GemmPlan plan_auto(int M, int N, int K, int kind) {
  GemmPlan pl{1, 128, 0};
  int cta_group = 0, block_n = 0;
  {
    // heuristic (measured on B200, profiles/r01_plan_probe.jsonl): small M streams weights with single-CTA tiles; up to
    // 256 rows one round of single-CTA 128 x 256 tiles; otherwise CTA pairs (256-row tiles halve the B traffic per flop)
    // with the tile width that minimises rounds x (k_tiles x w + o) -- w / o = per-k-tile / per-tile cost of one round in
    // us, fitted at K = 4096 and 8192: the 256-wide tile pays the single-buffered accumulator hand-off, the narrower ones
    // are bound by the tcgen05 dispatch rate (same per-k-tile cost for 192 and 128 columns).
    if (M <= 32 && decode_eligible(M, N, K, N, kind)) { cta_group = 1; block_n = 16; }
    else if (M <= 128) { cta_group = 1; block_n = 128; }
    else if (M <= 256 && N <= 4096) { cta_group = 1; block_n = 128; }
    else if (M <= 256 && ceil_div(M, 128) * ceil_div(N, 256) <= num_sms()) { cta_group = 1; block_n = 256; }
    else {
      cta_group = 2;
      const int64_t clusters = num_sms() / 2;
      const int64_t tm = ceil_div(M, 256);
      const double kt = (double)ceil_div(K, (kind == B200Q_KIND_MXF8 || kind == B200Q_KIND_MXF8_NN) ? 128 : 256);
      double best = -1.0;
      const int cands[3] = {256, 192, 128};
      const bool f8 = kind == B200Q_KIND_MXF8 || kind == B200Q_KIND_MXF8_NN;
      // MXFP8 (k-tile = 128 elements, twice the tensor time per byte): the narrower tiles are not dispatch-bound, their
      // per-k-tile cost follows the width (profiles/r02_s3_f8_probe.jsonl: 12.0 / 9.4 / 7.45 us per round at K = 4096)
      const double w4[3] = {0.36, 0.30, 0.30}, w8[3] = {0.3375, 0.267, 0.214}, o[3] = {1.2, 0.85, 0.6};
      const double* w = f8 ? w8 : w4;
      for (int i = 0; i < 3; ++i) {
        const int bn = cands[i];
        const int64_t rounds = ceil_div(tm * ceil_div(N, bn), clusters);
        const double cost = (double)rounds * (kt * w[i] + o[i]);
        if (best < 0 || cost < best) { best = cost; block_n = bn; }
      }
    }
  }
#ifdef B200Q_PROFILING
  if (env().gemm_hybrid && cta_group == 2 && hybrid_eligible(M, N, K, N, kind)) block_n = 448;
#endif
  pl.cta_group = cta_group;
  pl.block_n = block_n;
  // Measured (profiles/r01_notes.md): the peeled launch costs more than the saved part of a round (92.5 vs 89.4 us at
  // config 1 with (2,128) tail tiles, 97.1 us with (1,64)), because a tile's sequential k-loop, not the tile count, sets
  // the length of the last round.  Kept as an opt-in experiment: B200Q_TAIL_SPLIT=1.
  if (env().tail_split && cta_group == 2 && block_n == 256 && N % 8 == 0 && kind != B200Q_KIND_MXF8 && kind != B200Q_KIND_MXF8_NN) {
    // Wave quantisation: with T tiles over C CTA pairs the last round is only (T mod C)/C full.  When that round is
    // less than half full and peeling the LAST 256-column block of N saves a whole round, that block runs as a second
    // launch of narrower (2,128) tiles that fits in one wave (it starts as the main grid drains: PDL, disjoint D
    // columns, identical arithmetic -- every configuration is bit-identical).
    const int64_t clusters = num_sms() / 2;
    const int64_t tm = ceil_div(M, 256), tn = ceil_div(N, 256);
    const int64_t rounds = ceil_div(tm * tn, clusters);
    const int64_t waste = rounds * clusters - tm * tn;
    const int64_t rounds_main = tn > 1 ? ceil_div(tm * (tn - 1), clusters) : rounds;
    const int n_main = (int)(tn - 1) * 256;
    const int64_t tail_tiles = tm * ceil_div(N - n_main, 128);          // (2,128) tiles: half the k-loop latency of (2,256)
    if (tn > 1 && waste * 2 >= clusters && rounds_main < rounds && tail_tiles <= clusters) pl.n_main = n_main;
  }
  return pl;
}
}  // namespace b200q

using namespace b200q;

extern "C" int b200q_gemm_fp4_cfg(const void* A, const void* B, const void* SFA, const void* SFB,
                                  const float* alpha_dev, void* D_bf16, int M, int N, int K, int kind, int cta_group,
                                  int block_n, b200q_stream_t stream) {
  int rc = check_device_sm100();
  if (rc) return rc;
  const bool static_w = (kind & B200Q_GEMM_STATIC_WEIGHTS) != 0;
  kind &= ~B200Q_GEMM_STATIC_WEIGHTS;
  B200Q_REQUIRE(A && B && SFA && SFB && alpha_dev && D_bf16, "null pointer argument");
  B200Q_REQUIRE(kind == B200Q_KIND_MXF4 || kind == B200Q_KIND_NVF4 || kind == B200Q_KIND_MXF8 || kind == B200Q_KIND_MXF8_NN,
                "invalid kind %d", kind);
  B200Q_REQUIRE(kind != B200Q_KIND_MXF8_NN || M % 16 == 0, "nn: M (%d) must be a multiple of 16 (row pitch of A [K, M])", M);
  B200Q_REQUIRE(M > 0 && N > 0 && K > 0, "M, N, K must be positive (got %d, %d, %d)", M, N, K);
  B200Q_REQUIRE(K % 32 == 0, "K (%d) must be a multiple of 32", K);
  B200Q_REQUIRE((((uintptr_t)A | (uintptr_t)B | (uintptr_t)SFA | (uintptr_t)SFB) & 15) == 0,
                "A, B, SFA, SFB must be 16-byte aligned");
  B200Q_REQUIRE(((uintptr_t)D_bf16 & 15) == 0 || (N % 8) != 0, "D must be 16-byte aligned");
  const bool auto_cfg = (cta_group == 0 || block_n == 0);
  GemmPlan pl{cta_group, block_n, 0};
  if (auto_cfg) {
    pl = plan_auto(M, N, K, kind);
    cta_group = pl.cta_group;
    block_n = pl.block_n;
  }
  cudaStream_t s = (cudaStream_t)stream;
  auto run = [&](int cg, int bn, const void* Bp, const void* SFBp, void* Dp, int n_sub) -> int {
    if (kind == B200Q_KIND_NVF4) return dispatch_cfg<true, 0>(cg, bn, A, Bp, SFA, SFBp, alpha_dev, Dp, M, n_sub, K, N, s, static_w);
    if (kind == B200Q_KIND_MXF8) return dispatch_cfg<false, 1>(cg, bn, A, Bp, SFA, SFBp, alpha_dev, Dp, M, n_sub, K, N, s, static_w);
    if (kind == B200Q_KIND_MXF8_NN) return dispatch_cfg<false, 2>(cg, bn, A, Bp, SFA, SFBp, alpha_dev, Dp, M, n_sub, K, N, s, static_w);
    return dispatch_cfg<false, 0>(cg, bn, A, Bp, SFA, SFBp, alpha_dev, Dp, M, n_sub, K, N, s, static_w);
  };
  if (pl.n_main > 0) {
    const int group = kind == B200Q_KIND_NVF4 ? 16 : 32;
    const int64_t sf_col_blocks = ceil_div(ceil_div(K, group), 4);
    rc = run(2, 256, B, SFB, D_bf16, pl.n_main);
    if (rc) return rc;
    const uint8_t* Bt = (const uint8_t*)B + (int64_t)pl.n_main * (K / 2);
    const uint8_t* SFBt = (const uint8_t*)SFB + (int64_t)(pl.n_main / 128) * sf_col_blocks * 512;
    return run(2, 128, Bt, SFBt, (uint8_t*)D_bf16 + (int64_t)pl.n_main * 2, N - pl.n_main);
  }
  return run(cta_group, block_n, B, SFB, D_bf16, N);
}

// debug: copy the timeline of the last traced launch (B200Q_GEMM_DEBUG_FLAGS bit 24) into out[n] (synchronises)
extern "C" int b200q_debug_read_trace(unsigned long long* out, int n) {
  if (n > kTraceTiles * kTraceEvents) n = kTraceTiles * kTraceEvents;
  B200Q_CUDA(cudaDeviceSynchronize());
  B200Q_CUDA(cudaMemcpyFromSymbol(out, g_gemm_trace, sizeof(unsigned long long) * n));
  return 0;
}

extern "C" int b200q_debug_read_ktrace(unsigned long long* out, int n) {
  if (n > 8) n = 8;
  B200Q_CUDA(cudaDeviceSynchronize());
  B200Q_CUDA(cudaMemcpyFromSymbol(out, g_gemm_ktrace, sizeof(unsigned long long) * n));
  return 0;
}

extern "C" int b200q_debug_tmap_cache_stats(unsigned long long* hits, unsigned long long* misses) {
  B200Q_REQUIRE(hits && misses, "bad argument");
  *hits = g_tmap_hits;
  *misses = g_tmap_misses;
  g_tmap_hits = g_tmap_misses = 0;
  return 0;
}

extern "C" int b200q_gemm_fp4(const void* A, const void* B, const void* SFA, const void* SFB, const float* alpha_dev,
                              void* D_bf16, int M, int N, int K, int kind, b200q_stream_t stream) {
  return b200q_gemm_fp4_cfg(A, B, SFA, SFB, alpha_dev, D_bf16, M, N, K, kind, 0, 0, stream);
}

// ------------------------------------------------------------------ fused quantise + GEMM
namespace b200q {
// Measured on B200 (profiles/r01_notes.md, tools/fuse_probe.py): the fused kernel is bit-identical but NOT faster -- the
// part is power-limited under FP4 MMA load, the quantiser warps' work costs the tensor pipe about as much time as the
// standalone quantise kernel takes (104.8 us fused with the producer never waiting vs 103.7 us for the two launches,
// 93.8 us with the quantisers idle).  So the default is the two launches; B200Q_FUSE=1 opts into the single kernel.
static bool fusion_enabled() {
  return env().fuse != 0;
}
// the fused kernel needs: trusted Hadamard rotation, whole warp-tiles per row (K % 1024 == 0), TMA-store epilogue
// (N % 8 == 0), the CTA-pair plan (M > 256 or wide N) and FP4 operands
static bool fusable(int M, int N, int K, int had, int method, int kind) {
  if (!fusion_enabled()) return false;
  if (kind != B200Q_KIND_MXF4 && kind != B200Q_KIND_NVF4) return false;
  if (!(method & B200Q_ROT_TRUSTED_HADAMARD)) return false;
  if (K % 1024 != 0 || N % 8 != 0) return false;
  if ((int64_t)M * K / 1024 >= ((int64_t)1 << 31)) return false;
  (void)had;
  return plan_auto(M, N, K, kind).cta_group == 2;
}
}  // namespace b200q

extern "C" int64_t b200q_linear_fp4_workspace_bytes(int M) {
  if (M <= 0) return 0;
  return round_up((int64_t)(ceil_div(M, 256) + 2) * 4, 256);
}

extern "C" int b200q_linear_fp4_launches(int M, int N, int K, int had, int method, int kind) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  if (decode_fuse_eligible(M, N, K, had, method, kind)) return 1;
  if (fusable(M, N, K, had, method, kind)) return 1;
  return 1 + b200q_gemm_fp4_launches(M, N, K, kind);
}

extern "C" int b200q_linear_fp4(const void* x_bf16, const void* rot_bf16, void* xq_e2m1, void* x_sf_rowmajor,
                                void* x_sf_blocked, const void* Wq, const void* Wsf_blocked, const float* alpha_dev,
                                const float* global_scale_dev, void* D_bf16, void* ws, int M, int N, int K, int had,
                                int method, int kind, b200q_stream_t stream) {
  int rc = check_device_sm100();
  if (rc) return rc;
  B200Q_REQUIRE(kind == B200Q_KIND_MXF4 || kind == B200Q_KIND_NVF4, "invalid kind %d (MXF4 or NVF4)", kind);
  B200Q_REQUIRE(x_sf_blocked, "x_sf_blocked is required (the GEMM reads the blocked scales)");
  B200Q_REQUIRE(M > 0 && N > 0 && K > 0, "M, N, K must be positive (got %d, %d, %d)", M, N, K);
  const bool nv = kind == B200Q_KIND_NVF4;
  if (decode_fuse_eligible(M, N, K, had, method, kind)) {
    // decode (M <= 32): ONE launch -- every CTA of the weight-streaming kernel quantises the activations itself (gemm_decode.cu)
    B200Q_REQUIRE(Wq && Wsf_blocked && alpha_dev && D_bf16, "null pointer argument");
    const int mth = method & ~(B200Q_ROT_TRUSTED_HADAMARD | B200Q_ROT_GENERIC | B200Q_NV_SM100_CODES | B200Q_NV_ORACLE_CODES);
    B200Q_REQUIRE(mth == B200Q_METHOD_QUEST || mth == B200Q_METHOD_ABSMAX, "invalid method %d, must be quest (0) or abs_max (1)", mth);
    B200Q_REQUIRE(!nv || global_scale_dev, "global_scale must be a device pointer to one float");
    B200Q_REQUIRE((((uintptr_t)Wq | (uintptr_t)Wsf_blocked | (uintptr_t)D_bf16) & 15) == 0, "Wq, Wsf and D must be 16-byte aligned");
    QuantParams q;
    rc = fill_params(q, x_bf16, rot_bf16, xq_e2m1, x_sf_rowmajor, x_sf_blocked, (int64_t)M * K, K, had, nv ? 16 : 32);
    if (rc) return rc;
    note_sf_write(x_sf_rowmajor);
    q.gs = global_scale_dev;
    q.trust_hadamard = 1;
    q.nv_sm100_codes = (nv && had == 128 && mth == B200Q_METHOD_ABSMAX && !(method & B200Q_NV_ORACLE_CODES)) ? 1 : 0;
    return launch_gemm_decode_fused(q, had, mth, Wq, Wsf_blocked, alpha_dev, D_bf16, M, N, K, kind, (cudaStream_t)stream);
  }
  if (!ws || !fusable(M, N, K, had, method, kind)) {
    // two launches: the standalone quantiser, then the GEMM (programmatic dependent launch overlaps its prologue)
    rc = nv ? b200q_quantize_nv(x_bf16, rot_bf16, xq_e2m1, x_sf_rowmajor, x_sf_blocked, global_scale_dev, (int64_t)M * K, K,
                                had, method, stream)
            : b200q_quantize_mx(x_bf16, rot_bf16, xq_e2m1, x_sf_rowmajor, x_sf_blocked, nullptr, (int64_t)M * K, K, had,
                                method, stream);
    if (rc) return rc;
    return b200q_gemm_fp4(xq_e2m1, Wq, x_sf_blocked, Wsf_blocked, alpha_dev, D_bf16, M, N, K, kind | B200Q_GEMM_STATIC_WEIGHTS, stream);
  }
  B200Q_REQUIRE(Wq && Wsf_blocked && alpha_dev && D_bf16, "null pointer argument");
  B200Q_REQUIRE((((uintptr_t)xq_e2m1 | (uintptr_t)Wq | (uintptr_t)x_sf_blocked | (uintptr_t)Wsf_blocked | (uintptr_t)D_bf16) & 15) == 0,
                "xq, Wq, scale buffers and D must be 16-byte aligned");
  B200Q_REQUIRE(((uintptr_t)ws & 3) == 0, "workspace must be 4-byte aligned");
  const int m = method & ~(B200Q_ROT_TRUSTED_HADAMARD | B200Q_ROT_GENERIC | B200Q_NV_SM100_CODES | B200Q_NV_ORACLE_CODES);
  B200Q_REQUIRE(m == B200Q_METHOD_QUEST || m == B200Q_METHOD_ABSMAX, "invalid method %d, must be quest (0) or abs_max (1)", m);
  B200Q_REQUIRE(had == 32 || had == 64 || had == 128 || (nv && had == 16),
                nv ? "Unsupported rotation size %d; expected 16, 32, 64, or 128." : "Unsupported rotation size %d; expected 32, 64, or 128.", had);
  B200Q_REQUIRE(!nv || global_scale_dev, "global_scale must be a device pointer to one float");
  FuseParams fp = {};
  rc = fill_params(fp.q, x_bf16, rot_bf16, xq_e2m1, x_sf_rowmajor, x_sf_blocked, (int64_t)M * K, K, had, nv ? 16 : 32);
  if (rc) return rc;
  note_sf_write(x_sf_rowmajor);
  fp.q.gs = global_scale_dev;
  fp.q.trust_hadamard = 1;
  fp.q.nv_sm100_codes = (nv && had == 128 && m == B200Q_METHOD_ABSMAX && !(method & B200Q_NV_ORACLE_CODES)) ? 1 : 0;
  fp.ctr = (uint32_t*)ws;
  fp.tiles_per_row = (uint32_t)(K / 1024);
  fp.had = had;
  fp.method = m;
  const GemmPlan pl = plan_auto(M, N, K, kind);
  cudaStream_t s = (cudaStream_t)stream;
  const int qwarps = env().fuse_warps;
#define B200Q_FCASE(BNV, QW)                                                                                                  \
  if (pl.block_n == BNV && qwarps == QW)                                                                                      \
    return nv ? launch_gemm<2, BNV, true, 128, false, QW>(xq_e2m1, Wq, x_sf_blocked, Wsf_blocked, alpha_dev, D_bf16, M, N, K, N, s, true, &fp) \
              : launch_gemm<2, BNV, false, 128, false, QW>(xq_e2m1, Wq, x_sf_blocked, Wsf_blocked, alpha_dev, D_bf16, M, N, K, N, s, true, &fp);
  B200Q_FCASE(256, 4)
  B200Q_FCASE(192, 4)
  B200Q_FCASE(128, 4)
  B200Q_FCASE(256, 2)
#undef B200Q_FCASE
  set_error("no fused configuration for block_n=%d", pl.block_n);
  return B200Q_EINVAL;
}

extern "C" int b200q_gemm_fp4_plan(int M, int N, int K, int kind, int* cta_group, int* block_n) {
  B200Q_REQUIRE(M > 0 && N > 0 && K > 0 && cta_group && block_n, "bad argument");
  const GemmPlan pl = plan_auto(M, N, K, kind);
  *cta_group = pl.cta_group;
  *block_n = pl.block_n;
  return 0;
}

extern "C" int b200q_gemm_fp4_launches(int M, int N, int K, int kind) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return plan_auto(M, N, K, kind).n_main > 0 ? 2 : 1;
}
