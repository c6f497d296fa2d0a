// Weight-streaming block-scaled FP4 GEMM for decode-size batches (M <= 32 by default, 64-row instantiation on request), sm_100a.
//
// Same contract and arithmetic as gemm_fp4_kernel (D = bf16(alpha * (A.SFA)(B.SFB)^T), ONE fp32 accumulation chain per
// output over all of K in ascending k, alpha once, one RNE) -- replaces, for small M, the reference's 128 x 128 decode tile
// (qutlass/csrc/gemm.cu:195-203).  What differs is the mapping onto the tensor core, chosen from a measured timeline of the
// general kernel at M = 16 (profiles/r02_decode_probe_before.jsonl): there the mainloop is NOT bandwidth-bound, it costs
// ~480 cycles per k-tile on the one issuing thread: 4 scale copies + a commit + 4 MMAs of ~68 cycles each, because the tensor
// core fetches BOTH 128-row operands from shared memory (~64 B/clk) and the A operand is 7/8 padding.  Here:
//
//   * operands are SWAPPED: the 128 weight rows of a tile are the M = 128 operand ("A") of tcgen05.mma, the activations
//     the N = NP (16 / 32 / 64) operand ("B"): D^T[n, m] accumulates in 128 TMEM lanes x NP columns.  An MMA is 128 x NP x 64
//     instead of 128 x 128 x 64: only the weight operand (4 KB per MMA) is a full 128 rows.
//   * ALL of x (NP rows x K/2 bytes) and ALL of its scales are loaded ONCE per CTA into shared memory -- two TMA ops, issued by
//     an idle epilogue warp the moment the grid dependency resolves -- and the scales are copied to TMEM once (K/128 resp.
//     K/64 tcgen05.cp, under the latency of the first weight tiles): per k-tile the issuing thread does 2 (MX) / 4 (NV)
//     weight-scale copies + 4 MMAs + 1 commit (measured: 401 cycles per k-tile, 323-346 with no weight traffic at all).
//   * the ring holds only weights: 17-18 KB per stage, 8-10 k-tiles in flight per SM.
//   * epilogue: thread = output column n (TMEM lane), registers = the NP rows m; lane pairs swap one value per row pair so that
//     every lane stores 4 bytes: 64 contiguous bytes per warp and row -- D is tiny (M x N x 2 bytes).
// The products are the same numbers in the same k order as in the general kernel, so the result is bit-identical to it
// (tests/test_gpu_parity.py::test_decode_kernel_*).
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

#include <cuda.h>

namespace b200q {
using namespace ptx;

constexpr int kDecThreads = 192;          // warp 0 producer, warp 1 MMA issuer, warps 2-5 epilogue (one per TMEM lane quarter)
constexpr int kDecQuantWarps = 8;         // fused variant: warps 2-9 rotate + quantise the activations (2-5 then turn epilogue)
constexpr int kDecFuseThreads = 64 + 32 * kDecQuantWarps;
constexpr int kDecMaxGroups = 16;         // fused: groups of 4 k-tiles (1024 elements of a row = one quantiser warp-tile)
constexpr int kDecMaxStages = 12;
constexpr int kDecSmemBudget = 227 * 1024;
constexpr int kDecEarlyStages = 4;        // static weights: stages whose REAL loads are issued before the grid dependency
constexpr int kDecPaceFree = 3;           // paced streaming: the first stages are requested at once, the later ones on a clock

struct DecodeParams {
  const float* alpha;
  __nv_bfloat16* d;
  int M, N, K, ldd;
  int k_tiles;        // K / 256
  int tiles;          // ceil(N / 128)
  int stages;         // weight ring depth (host: what fits next to x)
  int static_weights;
  int pace_cycles;    // > 0: weight stage g is requested no earlier than (g - kDecPaceFree) * pace_cycles after set-up (see launch_decode_t)
  int flags;          // profiling builds only: 1 << 24 timeline of CTA 0, 1 << 22 weights loaded for the first ring only
};

// timeline of CTA 0 (profiling flag 1 << 24), clock64 unless noted: [0] entry, [1] globaltimer at entry, [2] set-up done,
// [3] producer past griddepcontrol.wait, [4] activations landed, [5] activation scales copied, [6] first weight stage
// landed, [7] last MMA issued, [8] accumulator complete, [9] stores issued, [10] exit, [11] globaltimer at exit
__device__ unsigned long long g_decode_trace[16];
__device__ __forceinline__ void dtrace(int flags, int ev, bool wall = false) {
  if ((flags & (1 << 24)) && blockIdx.x == 0) {
    unsigned long long t;
    if (wall) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    else t = (unsigned long long)clock64();
    g_decode_trace[ev] = t;
  }
}

// Fused variant (ONE launch for the decode step).  Every CTA needs ALL of the quantised activations, and
// quantising them redundantly in every CTA was measured 1.2 - 2.5x SLOWER than two launches (8 warps per SM are latency-bound
// at ~1.4 us per 1024-element warp-tile: profiles/r02_decode_probe2_v2.jsonl).  So the CTAs run in CLUSTERS OF 8 that share the
// work: CTA r of a cluster rotates + quantises every 8th warp-tile and writes the codes / scale bytes straight into the
// resident shared-memory tiles of ALL 8 CTAs (st.shared::cluster), then arrives (release.cluster) on each CTA's barrier of that
// group of 4 k-tiles; the MMA warps pick the groups up one by one.  At M = 16, K = 4096 that is ONE warp-tile per warp.
// Same code (tile_stage_unpack / tile_rotate_hadamard / chunk_quantise) as the standalone butterfly quantiser => same bytes.
// Cluster 0 also writes the codes and scales to global memory (the outputs of b200q_linear_fp4).
constexpr int kDecCluster = 8;
struct DecodeFuse {
  QuantParams q;
  int had, method;
};

template <bool kNV>
struct DecodeCfg {
  static constexpr int SFKB = kNV ? 4 : 2;                 // 512-B scale blocks per 128 rows per k-tile
  static constexpr int W_BYTES = 128 * 128;                // weight tile: 128 rows x 128 bytes (256 e2m1)
  static constexpr int WSF_BYTES = SFKB * 512;
  static constexpr int STAGE_BYTES = W_BYTES + WSF_BYTES;  // 17 / 18 KB: multiple of 1024 (128B-swizzle atoms stay aligned)
  static constexpr int BAR_BYTES = 1024;
  static_assert(STAGE_BYTES % 1024 == 0, "stage alignment");
};

// one quantiser warp's share: warp-tiles t = g * M + m (group-major, so group 0 -- k-tiles 0..3 of every row -- completes first)
template <int HAD, bool NV, int METHOD, int NP>
__device__ __forceinline__ void decode_quantise(const DecodeFuse& f, uint4* stage, int qw, int lane, uint32_t x_base, uint32_t xsf_base,
                                                uint32_t xq_bar0, int M, int k_groups, bool write_global) {
  // qw = (rank in cluster) + kDecCluster * (quantiser warp): this warp's first warp-tile; stride = all quantiser warps of the cluster
  constexpr int kStride = kDecCluster * kDecQuantWarps;
  const QuantParams& p = f.q;
  const float c_scale = __bfloat162float(p.rot[0]);
  float gs = 1.f, gs_rcp = 1.f;
  if constexpr (NV) {
    gs = *p.gs;
    gs_rcp = rcp_approx_ftz(gs);
  }
  const int n_tiles = k_groups * M;
  auto load = [&](int t, uint4 (&dst)[4]) {
    const int g = t / M, m = t - g * M;
    const int64_t gt = (int64_t)m * k_groups + g;           // warp-tile index in x (row-major [M, K], 1024 elements each)
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = (t < n_tiles) ? __ldg(p.x + (gt * 128 + i * 32 + lane)) : make_uint4(0, 0, 0, 0);
  };
  uint4 nxt[4];
  load(qw, nxt);
  for (int t = qw; t < n_tiles; t += kStride) {
    uint4 ld[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld[i] = nxt[i];
    load(t + kStride, nxt);
    float v[32];
    tile_stage_unpack(ld, stage, lane, v);
    tile_rotate_hadamard<HAD>(v, c_scale);
    uint32_t out[4], sf_bytes, mask_word;
    chunk_quantise<NV, METHOD, false>(v, gs, gs_rcp, out, sf_bytes, mask_word, p.nv_sm100_codes != 0);
    const int g = t / M, m = t - g * M;
    {
      // codes: lane L owns 32 elements = one 16-byte piece of row m; k-tile g*4 + L/8, piece L%8, 128B swizzle (piece ^= row % 8)
      const int kt = g * 4 + (lane >> 3), j = lane & 7;
      const uint32_t addr = x_base + (uint32_t)kt * (NP * 128u) + (uint32_t)(m >> 3) * 1024u + (uint32_t)(m & 7) * 128u +
                            (uint32_t)((j ^ (m & 7)) << 4);
      // scales: blocked layout of the (single) 128-row block: block c/4, byte (m % 32) * 16 + (m / 32) * 4 + c % 4
      const int c = NV ? 2 * (g * 32 + lane) : (g * 32 + lane);
      const uint32_t saddr = xsf_base + (uint32_t)(c >> 2) * 512u + (uint32_t)(m & 31) * 16u + (uint32_t)(m >> 5) * 4u + (uint32_t)(c & 3);
      // the 4 scale bytes of one blocked cell sit in 4 (MX) / 2 (NV) neighbouring lanes: gather them into the cell's first lane
      uint32_t cell = sf_bytes;
      if constexpr (NV) {
        cell |= __shfl_down_sync(0xffffffffu, sf_bytes, 1) << 16;
      } else {
        cell |= __shfl_down_sync(0xffffffffu, sf_bytes, 1) << 8;
        cell |= __shfl_down_sync(0xffffffffu, sf_bytes, 2) << 16;
        cell |= __shfl_down_sync(0xffffffffu, sf_bytes, 3) << 24;
      }
      const bool cell_owner = NV ? ((lane & 1) == 0) : ((lane & 3) == 0);
      // Into the resident tiles of EVERY CTA of the cluster (identical shared-memory layout in all of them) with st.async:
      // each store's bytes are counted on THAT CTA's barrier of the group (armed with expect_tx at set-up), so the producer
      // needs neither a proxy fence nor a cluster-scope release arrive -- those cost ~3 us in the first cluster version.
      const uint32_t gbar = xq_bar0 + 8u * (uint32_t)g;
#pragma unroll
      for (uint32_t d = 0; d < (uint32_t)kDecCluster; ++d) {
        const uint32_t rbar = mapa(gbar, d);
        st_async_v4(mapa(addr, d), out[0], out[1], out[2], out[3], rbar);
        if (cell_owner) st_async_b32(mapa(saddr, d), cell, rbar);
      }
    }
    if (write_global) chunk_store<NV, false>(p, ((int64_t)m * k_groups + g) * 32 + lane, out, sf_bytes, 0u);
  }
}

template <bool NV, int NP>
__device__ __forceinline__ void decode_quantise_role(const DecodeFuse& f, uint4* stage, int qw, int lane, uint32_t x_base,
                                                     uint32_t xsf_base, uint32_t xq_bar0, int M, int k_groups, bool wg) {
#define B200Q_DQ(H)                                                                                                                     \
  case H:                                                                                                                                \
    if (f.method == B200Q_METHOD_QUEST) decode_quantise<H, NV, B200Q_METHOD_QUEST, NP>(f, stage, qw, lane, x_base, xsf_base, xq_bar0, M, k_groups, wg); \
    else decode_quantise<H, NV, B200Q_METHOD_ABSMAX, NP>(f, stage, qw, lane, x_base, xsf_base, xq_bar0, M, k_groups, wg);                \
    break;
  switch (f.had) {
    B200Q_DQ(128)
    B200Q_DQ(64)
    B200Q_DQ(32)
    case 16:
      if constexpr (NV) {
        if (f.method == B200Q_METHOD_QUEST) decode_quantise<16, NV, B200Q_METHOD_QUEST, NP>(f, stage, qw, lane, x_base, xsf_base, xq_bar0, M, k_groups, wg);
        else decode_quantise<16, NV, B200Q_METHOD_ABSMAX, NP>(f, stage, qw, lane, x_base, xsf_base, xq_bar0, M, k_groups, wg);
      }
      break;
  }
#undef B200Q_DQ
}

template <bool kNV, int NP, bool kFuse>
__global__ void __launch_bounds__(kFuse ? kDecFuseThreads : kDecThreads, 1)
gemm_fp4_decode_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                       const __grid_constant__ CUtensorMap tmap_sfx, const __grid_constant__ CUtensorMap tmap_sfw,
                       const DecodeParams p, const DecodeFuse fq) {
  using Cfg = DecodeCfg<kNV>;
  constexpr int SFKB = Cfg::SFKB;
  constexpr int ACC = 2;
  static_assert(NP == 16 || NP == 32 || NP == 64, "activation rows per MMA");
  static_assert(!kFuse || NP <= 32, "the one-launch variant covers M <= 32");
  const int STAGES = p.stages;

  extern __shared__ uint8_t dec_smem_raw[];
  const uint32_t smem_base = (smem_u32(dec_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = dec_smem_raw + (smem_base - smem_u32(dec_smem_raw));
  const uint32_t x_bytes = (uint32_t)p.k_tiles * NP * 128u;                // resident activations, one 128B-swizzled tile per k-tile
  const uint32_t xsf_bytes = (uint32_t)p.k_tiles * SFKB * 512u;            // resident activation scales (blocked layout)
  const uint32_t x_base = smem_base;
  const uint32_t xsf_base = x_base + x_bytes;
  const uint32_t qstage_bytes = kFuse ? kDecQuantWarps * 2048u : 0u;     // fused: 2 KB of swizzled staging per quantiser warp
  const uint32_t ring_base = (xsf_base + xsf_bytes + qstage_bytes + 1023u) & ~1023u;
  const uint32_t bar_base = ring_base + (uint32_t)STAGES * Cfg::STAGE_BYTES;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kDecMaxStages + s); };
  const uint32_t x_bar = bar_base + 8u * (2 * kDecMaxStages);
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kDecMaxStages + 1 + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kDecMaxStages + 1 + ACC + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kDecMaxStages + 1 + 2 * ACC);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (bar_base - smem_base) + 8 * (2 * kDecMaxStages + 1 + 2 * ACC));
  const uint32_t xq_bar0 = bar_base + 8u * (2 * kDecMaxStages + 2 + 2 * ACC);     // fused: one barrier per group of 4 k-tiles
  const int k_groups = p.k_tiles >> 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  if (threadIdx.x == 0) { dtrace(B200Q_FLAGS(p), 0); dtrace(B200Q_FLAGS(p), 1, true); }

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_w);
    prefetch_tensormap(&tmap_sfw);
    if constexpr (!kFuse) {
      prefetch_tensormap(&tmap_x);
      prefetch_tensormap(&tmap_sfx);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(x_bar, 1);
    if constexpr (kFuse) {
      // one barrier per group of 4 k-tiles, armed once: M rows x (512 B of codes + 32 / 64 B of scales) arrive by st.async from
      // the 8 CTAs of the cluster
      for (int g = 0; g < k_groups; ++g) {
        mbar_init(xq_bar0 + 8u * g, 1);
        mbar_arrive_expect_tx(xq_bar0 + 8u * g, (uint32_t)p.M * (512u + (kNV ? 64u : 32u)));
      }
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);            // the four epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  if constexpr (kFuse) cluster_sync();      // peers write into our tiles and arrive on our barriers: they must be initialised first
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_gen, 0);
  const uint32_t tmem_wsf = tmem_base + ACC * NP;              // weight scales of the current k-tile (SFKB blocks x 4 columns)
  const uint32_t tmem_xsf = tmem_wsf + SFKB * 4;               // ALL activation scales: k_tiles x SFKB blocks x 4 columns
  if (threadIdx.x == 0) dtrace(B200Q_FLAGS(p), 2);
  // Paced streaming (p.pace_cycles > 0).  The memory system serves everything that is outstanding at once, interleaved:
  // a whole ring requested at set-up lands as ONE bunch (measured: first stage 2.7-3.6 us after entry with 8-10 stages in
  // flight on 112 SMs), the refills only start then, and HBM idles for a memory latency between the bunches (7.6 us for
  // 31.2 MB).  Requesting stage g no earlier than (g - kDecPaceFree) * pace_cycles after set-up -- pace = this CTA's share of
  // the HBM rate -- keeps the requests of all CTAs in k order and the stream continuous.  Same loads, same bytes.
  const long long pace_t0 = clock64();
  // Only the CTA's FIRST tile is paced (and prefetched): that is where all CTAs start together; later tiles stream in steady
  // state, paced by the ring itself -- a clock there could only throttle (and warp 3 has epilogues to run).
  auto pace_gate = [&](int g) {
    if (p.pace_cycles > 0 && g > kDecPaceFree && g < p.k_tiles) {
      const long long due = (long long)(g - kDecPaceFree) * p.pace_cycles;
      while (clock64() - pace_t0 < due) {}
    }
  };

  int my_tiles = 0;
  for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) ++my_tiles;
  const int total_kt = my_tiles * p.k_tiles;

  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool elected = elect_one();
    auto load_w = [&](int stage, int tile, int kt) {
      const uint32_t sb = ring_base + (uint32_t)stage * Cfg::STAGE_BYTES;
      mbar_arrive_expect_tx(full_bar(stage), (uint32_t)Cfg::STAGE_BYTES);
      tma_load_2d<1>(sb, &tmap_w, full_bar(stage), kt * 128, tile * 128);
      tma_load_3d<1>(sb + Cfg::W_BYTES, &tmap_sfw, full_bar(stage), 0, kt * SFKB, tile);
    };
    auto prefetch_w = [&](int tile, int kt) {
      tma_prefetch_2d(&tmap_w, kt * 128, tile * 128);
      tma_prefetch_3d(&tmap_sfw, 0, kt * SFKB, tile);
    };
    const int pre = total_kt < STAGES ? total_kt : STAGES;
    // Before the grid dependency resolves only the weights may be touched, and only for real if the caller vouched for
    // them (B200Q_GEMM_STATIC_WEIGHTS) -- and then only a few stages, so that the activation loads issued right after the
    // wait are not queued behind a whole ring of weight traffic; everything else is an L2 prefetch (always safe).
    const int early = p.static_weights ? (pre < kDecEarlyStages ? pre : kDecEarlyStages) : 0;
    {
      int tile = blockIdx.x, kt = 0;
      for (int g = 0; g < pre; ++g) {
        if (elected) {
          if (g < early) load_w(g, tile, kt);
          else if (p.pace_cycles <= 0 || g >= p.k_tiles) prefetch_w(tile, kt);   // paced: warp 3 prefetches the first tile's k-tiles on the clock
        }
        if (++kt == p.k_tiles) { kt = 0; tile += gridDim.x; }
      }
    }
    pdl_wait();
    if (lane == 0) dtrace(B200Q_FLAGS(p), 3);
    int tile = blockIdx.x, kt = 0;
    for (int g = 0; g < pre; ++g) {
      if (g >= early) {
        pace_gate(g);
        if (elected) load_w(g, tile, kt);
      }
      if (++kt == p.k_tiles) { kt = 0; tile += gridDim.x; }
    }
    __syncwarp();
    int stage = (pre == STAGES) ? 0 : pre;
    uint32_t phase = (pre == STAGES) ? 1 : 0;
    const bool ring_only = (B200Q_FLAGS(p) & (1 << 22)) != 0;      // profiling: later k-tiles reuse what the first ring loaded
    for (int g = pre; g < total_kt; ++g) {
      mbar_wait(empty_bar(stage), phase ^ 1, 1);
      pace_gate(g);
      if (elected) {
        if (ring_only) mbar_arrive(full_bar(stage));
        else load_w(stage, tile, kt);
      }
      __syncwarp();
      if (++kt == p.k_tiles) { kt = 0; tile += gridDim.x; }
      if (++stage == STAGES) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool elected = elect_one();
    constexpr uint32_t idesc_base = make_idesc_fp4(128, NP, !kNV);
    constexpr uint32_t kDescHiAB = (1024u >> 4) | (1u << 14) | (kLayoutSw128 << 29);   // SBO 1024 B, version 1, 128B swizzle
    constexpr uint32_t kDescHiSF = (128u >> 4) | (1u << 14);                            // SBO 128 B, version 1, no swizzle
    auto mk = [](uint32_t lo, uint32_t hi) {
      uint64_t d;
      asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
      return d;
    };
    const uint32_t x_lo0 = ((x_base & 0x3FFFFu) >> 4) | (1u << 16);
    const uint32_t xsf_lo0 = (xsf_base & 0x3FFFFu) >> 4;
    const uint32_t ring_lo0 = ((ring_base & 0x3FFFFu) >> 4);
    if constexpr (!kFuse) {
      // activation scales -> TMEM, once, all up front.  (Measured, profiles/r02_s3_decode_interleave_*.jsonl: copying only the first
      // k-tile's blocks here and the rest one k-tile ahead inside the loop is SLOWER -- M = 16, L2-resident weights: 5.90 -> 6.39
      // us; the tensor pipe runs cp / mma in order and the extra copies lengthen every k-tile of the issue-bound loop.)
      mbar_wait(x_bar, 0, 2);
      tc_fence_after();
      if (lane == 0) dtrace(B200Q_FLAGS(p), 4);
      const int nblk = p.k_tiles * SFKB;
      for (int c = 0; c < nblk; ++c) {
        if (elected) tmem_cp_32x128b_warpx4<1>(tmem_xsf + (uint32_t)c * 4u, mk(xsf_lo0 + (uint32_t)c * 32u, kDescHiSF));
      }
      __syncwarp();
      if (lane == 0) dtrace(B200Q_FLAGS(p), 5);
    }
    bool first_tile = true;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (B200Q_FLAGS(p) & (1 << 24)) {
      mbar_wait(full_bar(0), 0, 9);
      if (lane == 0) dtrace(B200Q_FLAGS(p), 6);
    }
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1, 3);
      tc_fence_after();
      const uint32_t tmem_acc = tmem_base + (uint32_t)acc * NP;
      for (int kt = 0; kt < p.k_tiles; ++kt) {
        if constexpr (kFuse) {
          if (first_tile && (kt & 3) == 0) {
            // fused: group kt/4 of the activations (4 k-tiles of every row) has been quantised into shared memory by the
            // quantiser warps of THIS CTA -> its scale blocks go to TMEM now
            mbar_wait<true>(xq_bar0 + 8u * (uint32_t)(kt >> 2), 0, 7);      // acquire at cluster scope: the writers are 8 CTAs
            fence_proxy_async_smem();                                       // ... and the readers are the tensor core's async proxy
            tc_fence_after();
            if (kt == 0 && lane == 0) dtrace(B200Q_FLAGS(p), 4);
            for (int c = kt * SFKB; c < (kt + 4) * SFKB; ++c) {
              if (elected) tmem_cp_32x128b_warpx4<1>(tmem_xsf + (uint32_t)c * 4u, mk(xsf_lo0 + (uint32_t)c * 32u, kDescHiSF));
            }
            __syncwarp();
          }
        }
        mbar_wait_spin(full_bar(stage), phase);       // (a lost arrival is caught by the timed waits of the producer / epilogue)
        tc_fence_after();
        const uint32_t w_lo = (ring_lo0 + (uint32_t)stage * (Cfg::STAGE_BYTES >> 4)) | (1u << 16);
        const uint32_t wsf_lo = ring_lo0 + (uint32_t)stage * (Cfg::STAGE_BYTES >> 4) + (Cfg::W_BYTES >> 4);
        const uint32_t x_lo = x_lo0 + (uint32_t)kt * ((NP * 128u) >> 4);
        const uint32_t txsf = tmem_xsf + (uint32_t)kt * (SFKB * 4u);
        if (elected) {
#pragma unroll
          for (int kb = 0; kb < 4; ++kb) {
            // scales per MMA (K = 64): MXF4 2 bytes of a 4-byte cell (sf_id 0 / 2), NVF4 the whole cell
            const uint32_t chunk = kNV ? (uint32_t)kb : (uint32_t)(kb >> 1);
            if (kNV || (kb & 1) == 0) tmem_cp_32x128b_warpx4<1>(tmem_wsf + chunk * 4u, mk(wsf_lo + chunk * 32u, kDescHiSF));
            const uint32_t sf_id = kNV ? 0u : (uint32_t)((kb & 1) * 2);
            // A operand = weights (scales: tmem_wsf), B operand = activations (scales: the resident block of this k-tile)
            mma_fp4_block_scaled<1, kNV, false>(tmem_acc, mk(w_lo + kb * 2u, kDescHiAB), mk(x_lo + kb * 2u, kDescHiAB),
                                                idesc_base | (sf_id << 4) | (sf_id << 29), tmem_wsf + chunk * 4u, txsf + chunk * 4u,
                                                (kb > 0) ? 1u : (kt > 0 ? 1u : 0u));
          }
          tc_commit<1>(empty_bar(stage));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (elected) tc_commit<1>(tfull_bar(acc));
      __syncwarp();
      first_tile = false;
      if (lane == 0) dtrace(B200Q_FLAGS(p), 7);
      if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5): thread = output column, registers = the rows =====================
    const int q = warp & 3;                       // TMEM lane quarter this warp may access
    int acc = 0;
    uint32_t acc_phase = 0;
    if (!kFuse && warp == 3 && p.pace_cycles > 0) {
      // the weight stream into L2, on the clock, from kernel entry on -- also while the predecessor kernel is still running
      // (an L2 prefetch never makes stale data visible); the ring loads behind it hit L2 or join the fill in flight
      if (elect_one()) {
        const int tile = blockIdx.x;
        for (int kt = 0; kt < p.k_tiles && kt < total_kt; ++kt) {
          pace_gate(kt);
          tma_prefetch_2d(&tmap_w, kt * 128, tile * 128);
          tma_prefetch_3d(&tmap_sfw, 0, kt * SFKB, tile);
        }
      }
      __syncwarp();
    }
    pdl_wait();                                   // x comes from the predecessor kernel; D may still be read by it
    if constexpr (kFuse) {
      // warps 2..9: rotate + quantise this CTA's share of x into the resident tiles of the whole cluster; cluster 0 also
      // writes the outputs of the quantiser to global memory
      const int qwl = warp - 2;
      uint4* qstage = reinterpret_cast<uint4*>(smem_gen + (xsf_base - smem_base) + xsf_bytes) + qwl * 128;
      const int qw = (int)cluster_ctarank() + kDecCluster * qwl;
      decode_quantise_role<kNV, NP>(fq, qstage, qw, lane, x_base, xsf_base, xq_bar0, p.M, k_groups, blockIdx.x < kDecCluster);
      // (after the quantisation: the whole cluster waits for CTA 0's share)
      if (blockIdx.x == 0) zero_fill_sf_padding(fq.q, (int64_t)threadIdx.x - 64, (int64_t)kDecQuantWarps * 32);
      if (warp >= 6) goto done;                    // the four extra warps have no epilogue duty
    }
    if (!kFuse && warp == 2) {
      // The activations and their scales, resident for the whole kernel: TWO TMA ops, issued by this otherwise idle warp the
      // moment the grid dependency resolves -- the producer warp is busy issuing a ring of weight loads at that time, and
      // behind those (measured, profiles/r02_decode_probe2_v1.jsonl) x landed 1.9 k cycles later than it had to.
      if (elect_one()) {
        mbar_arrive_expect_tx(x_bar, x_bytes + xsf_bytes);
        tma_load_3d<1>(xsf_base, &tmap_sfx, x_bar, 0, 0, 0);      // all scale blocks of the (one) 128-row block
        tma_load_3d<1>(x_base, &tmap_x, x_bar, 0, 0, 0);          // all k-tiles: {128 B, NP rows, k_tiles} in ONE box
      }
      __syncwarp();
    }
    const float alpha = __ldg(p.alpha);
    for (int t = blockIdx.x; t < p.tiles; t += gridDim.x) {
      mbar_wait(tfull_bar(acc), acc_phase, 4);
      tc_fence_after();
      if (warp == 2 && lane == 0) dtrace(B200Q_FLAGS(p), 8);
      uint32_t r[NP];
      const uint32_t taddr = tmem_base + (uint32_t)acc * NP + ((uint32_t)(q * 32) << 16);
      if constexpr (NP == 16) tmem_ld_32x32b_x16(taddr, r);
      else {
#pragma unroll
        for (int j = 0; j < NP / 32; ++j) tmem_ld_32x32b_x32(taddr + j * 32, r + j * 32);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      const int n = t * 128 + q * 32 + lane;
      if ((p.ldd & 1) == 0) {
        // rows in pairs: the even lane of a lane pair stores (n, n+1) of row m, the odd lane (n-1, n) of row m+1 -- one 4-byte
        // store per lane and row pair, 64 contiguous bytes per warp and row
        const bool odd = lane & 1;
        const int n_pair = n & ~1;
#pragma unroll
        for (int m = 0; m < NP; m += 2) {
          const uint32_t mine0 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(__uint_as_float(r[m]) * alpha));
          const uint32_t mine1 = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(__uint_as_float(r[m + 1]) * alpha));
          const uint32_t send = odd ? mine0 : mine1;                       // what the partner's row needs from me
          const uint32_t got = __shfl_xor_sync(0xffffffffu, send, 1);
          const uint32_t word = odd ? (got | (mine1 << 16)) : (mine0 | (got << 16));
          const int row = m + (odd ? 1 : 0);
          if (row < p.M && n_pair < p.N) {
            uint16_t* dst = reinterpret_cast<uint16_t*>(p.d) + (int64_t)row * p.ldd + n_pair;
            if (n_pair + 1 < p.N) *reinterpret_cast<uint32_t*>(dst) = word;
            else dst[0] = (uint16_t)(word & 0xffffu);
          }
        }
      } else if (n < p.N) {
        uint16_t* dcol = reinterpret_cast<uint16_t*>(p.d) + n;
#pragma unroll
        for (int m = 0; m < NP; ++m) {
          if (m < p.M) dcol[(int64_t)m * p.ldd] = __bfloat16_as_ushort(__float2bfloat16_rn(__uint_as_float(r[m]) * alpha));
        }
      }
      if (warp == 2 && lane == 0) dtrace(B200Q_FLAGS(p), 9);
      if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
    }
  }

done:
  tc_fence_before();
  if constexpr (kFuse) cluster_sync();      // no CTA may exit while a peer can still write into its shared memory
  else __syncthreads();
  if (threadIdx.x == 0) { dtrace(B200Q_FLAGS(p), 10); dtrace(B200Q_FLAGS(p), 11, true); }
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ host side
static int decode_stages(int M, int K, bool nv, int* np_out, bool fuse = false) {
  const int np = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  const int k_tiles = K / 256;
  const int sfkb = nv ? 4 : 2;
  const int64_t resident = (int64_t)k_tiles * np * 128 + (int64_t)k_tiles * sfkb * 512 + (fuse ? kDecQuantWarps * 2048 : 0);
  const int stage = 128 * 128 + sfkb * 512;
  const int64_t room = (int64_t)kDecSmemBudget - 1024 /*alignment*/ - 1024 /*ring alignment*/ - 1024 /*barriers*/ - resident;
  int stages = (int)(room / stage);
  if (stages > kDecMaxStages) stages = kDecMaxStages;
  if (np_out) *np_out = np;
  return stages;
}

bool decode_eligible(int M, int N, int K, int ldd, int kind) {
  if (kind != B200Q_KIND_MXF4 && kind != B200Q_KIND_NVF4) return false;
  if (M < 1 || M > 64 || N < 1 || K < 256 || K % 256 != 0 || ldd < N) return false;
  const bool nv = kind == B200Q_KIND_NVF4;
  const int k_tiles = K / 256, sfkb = nv ? 4 : 2;
  const int np = M <= 16 ? 16 : (M <= 32 ? 32 : 64);
  if (2 * np + sfkb * 4 + k_tiles * sfkb * 4 > 512) return false;      // TMEM: accumulators + weight scales + ALL activation scales
  if (k_tiles * sfkb > 256) return false;                                  // one tensor-map box holds all activation scale blocks
  if (k_tiles > 256) return false;                                         // ... and one box all k-tiles of x
  return decode_stages(M, K, nv, nullptr) >= 4;
}

bool decode_fuse_eligible(int M, int N, int K, int had, int method, int kind) {
  // Measured on B200, quantise + GEMM per step under graph replay, M = 1 / 16 / 32 (profiles/r02_decode_probe2_v{2,3}.jsonl):
  //   two launches (default)                              8.6 /  8.4 /  8.9 us
  //   one launch, every CTA quantising ALL of x itself    9.8 / 14.4 / 22.5 us
  //   one launch, clusters of 8 sharing the quantisation 11.5 / 12.6 / 13.8 us  (timeline: the first group of quantised
  //     activations reaches the MMA warp 6.5 us after kernel entry -- remote stores + fence.proxy.async + release.cluster
  //     arrives cost more than the kernel boundary they replace)
  // Both one-launch forms are bit-identical to the two launches; neither is faster, so the single launch is opt-in.
  if (!env().fuse_decode) return false;
  if (M > 32 || !decode_eligible(M, N, K, N, kind)) return false;
  if (!(method & B200Q_ROT_TRUSTED_HADAMARD)) return false;              // in-register butterflies only
  const bool nv = kind == B200Q_KIND_NVF4;
  if (!(had == 32 || had == 64 || had == 128 || (nv && had == 16))) return false;
  if (K % 1024 != 0 || K / 1024 > kDecMaxGroups) return false;           // whole quantiser warp-tiles per row, one barrier per group
  return decode_stages(M, K, nv, nullptr, true) >= 4;
}

template <bool kNV, int NP, bool kFuse = false>
static int launch_decode_t(const void* A, const void* B, const void* SFA, const void* SFB, const float* alpha, void* D, int M, int N,
                           int K, int ldd, bool static_w, cudaStream_t stream, const DecodeFuse* fuse = nullptr) {
  using Cfg = DecodeCfg<kNV>;
  auto kern = gemm_fp4_decode_kernel<kNV, NP, kFuse>;
  DecodeParams p;
  p.alpha = alpha;
  p.d = (__nv_bfloat16*)D;
  p.M = M; p.N = N; p.K = K; p.ldd = ldd;
  p.k_tiles = K / 256;
  p.tiles = (int)ceil_div(N, 128);
  p.stages = decode_stages(M, K, kNV, nullptr, kFuse);
  p.static_weights = static_w ? 1 : 0;
  // Paced weight streaming (kernel comment at pace_gate).  Measured (profiles/r02_s3_decode_pace_{mx,nv}.jsonl, N = 14336,
  // K = 4096, graph replay, weights from HBM): default launches 8.14 / 7.83 / 8.51 us -> 7.97 / 7.75 / 8.12 (M = 1 / 16 / 32) at the
  // pace below, NVFP4 8.63 / 8.41 / 8.99 -> 8.31 / 8.07 / 8.88; L2-resident weights unchanged (the MMA loop is slower than the
  // clock).  Static-weights launches gain more from HBM (7.54 -> 7.06 at M = 16) but lose 0.3 us at M = 1 with L2-resident
  // weights, so they stay unpaced unless B200Q_DECODE_PACE says otherwise.  pace = the CTA's share of ~3600 B per SM cycle
  // (6.8 TB/s at 1.9 GHz): 540 cycles for 112 CTAs x 17 KB stages.  B200Q_DECODE_PACE=0 switches it off, > 0 forces cycles.
  p.pace_cycles = 0;
  if (!kFuse) {
    const int sw = env().decode_pace;
    int tiles_now = (int)ceil_div(N, 128);
    if (tiles_now > num_sms()) tiles_now = num_sms();
    if (sw > 0) p.pace_cycles = sw;
    else if (sw < 0 && !static_w) p.pace_cycles = (int)((int64_t)Cfg::STAGE_BYTES * tiles_now / 3600);
  }
  p.flags = env().gemm_flags;
  const int smem = 1024 + p.k_tiles * NP * 128 + p.k_tiles * Cfg::SFKB * 512 + (kFuse ? kDecQuantWarps * 2048 : 0) + 1024 +
                   p.stages * Cfg::STAGE_BYTES + Cfg::BAR_BYTES;
  static std::atomic<unsigned long long> smem_attr_done{0};
  if (int rc_attr = ensure_dynamic_smem(kern, kDecSmemBudget, smem_attr_done)) return rc_attr;
  const int group = kNV ? 16 : 32;
  const int64_t sf_col_blocks = ceil_div(ceil_div(K, group), 4);
  CUtensorMap tx, tw, tsx, tsw;
  int rc;
  if ((rc = make_operand_tmap(&tw, B, N, K / 2, 128, "W (decode)"))) return rc;
  if ((rc = make_sf_tmap(&tsw, SFB, ceil_div(N, 128), sf_col_blocks, Cfg::SFKB, 1, "SFW (decode)"))) return rc;
  if constexpr (kFuse) {
    tx = tw;        // the fused kernel quantises x itself: the activation maps are unused
    tsx = tsw;
  } else {
    if ((rc = make_operand_ktile_tmap(&tx, A, M, K / 2, NP, "x (decode)"))) return rc;
    if ((rc = make_sf_tmap(&tsx, SFA, ceil_div(M, 128), sf_col_blocks, (int)sf_col_blocks, 1, "SFx (decode)"))) return rc;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kFuse ? kDecFuseThreads : kDecThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[2];
  cfg.attrs = attrs;
  int ctas = num_sms();
  if (ctas > p.tiles) ctas = p.tiles;
  int n_attr = 0;
  if constexpr (kFuse) {
    // clusters of 8 (one GPC each): as many as are co-resident, every cluster complete (CTAs without a tile still quantise)
    attrs[n_attr].id = cudaLaunchAttributeClusterDimension;
    attrs[n_attr].val.clusterDim.x = kDecCluster;
    attrs[n_attr].val.clusterDim.y = 1;
    attrs[n_attr].val.clusterDim.z = 1;
    ++n_attr;
    static std::atomic<int> max_clusters[64];
    std::atomic<int>& mc_a = max_clusters[current_device() & 63];
    int mc = mc_a.load(std::memory_order_acquire);
    if (mc == 0) {
      cfg.gridDim = dim3((unsigned)(num_sms() / kDecCluster * kDecCluster));
      cfg.numAttrs = n_attr;
      int n = 0;
      B200Q_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &cfg));
      mc = n > 0 ? n : 1;
      mc_a.store(mc, std::memory_order_release);
    }
    int clusters = (int)ceil_div(p.tiles, kDecCluster);
    if (clusters > mc) clusters = mc;
    ctas = clusters * kDecCluster;
  }
  cfg.gridDim = dim3((unsigned)ctas);
  if (env().no_pdl != 1) {
    attrs[n_attr].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attrs[n_attr].val.programmaticStreamSerializationAllowed = 1;
    ++n_attr;
  }
  cfg.numAttrs = n_attr;
  DecodeFuse fz = {};
  if (fuse) fz = *fuse;
  B200Q_CUDA(cudaLaunchKernelEx(&cfg, kern, tx, tw, tsx, tsw, p, fz));
  return 0;
}

// the whole decode step in ONE launch: q describes the quantiser's inputs / outputs (fill_params), had / method as in b200q.h
int launch_gemm_decode_fused(const QuantParams& q, int had, int method, const void* B, const void* SFB, const float* alpha, void* D,
                             int M, int N, int K, int kind, cudaStream_t stream) {
  DecodeFuse f;
  f.q = q;
  f.had = had;
  f.method = method;
  const bool nv = kind == B200Q_KIND_NVF4;
  // b200q_linear_fp4 takes pre-quantised weights by contract: static
  if (M <= 16) return nv ? launch_decode_t<true, 16, true>(nullptr, B, nullptr, SFB, alpha, D, M, N, K, N, true, stream, &f)
                         : launch_decode_t<false, 16, true>(nullptr, B, nullptr, SFB, alpha, D, M, N, K, N, true, stream, &f);
  return nv ? launch_decode_t<true, 32, true>(nullptr, B, nullptr, SFB, alpha, D, M, N, K, N, true, stream, &f)
            : launch_decode_t<false, 32, true>(nullptr, B, nullptr, SFB, alpha, D, M, N, K, N, true, stream, &f);
}

int launch_gemm_decode(const void* A, const void* B, const void* SFA, const void* SFB, const float* alpha, void* D, int M, int N,
                       int K, int ldd, int kind, bool static_w, cudaStream_t stream) {
  if (!decode_eligible(M, N, K, ldd, kind)) {
    set_error("the decode kernel needs an FP4 kind, 1 <= M <= 64, K %% 256 == 0 and K small enough for resident activations "
              "(M=%d N=%d K=%d kind=%d)", M, N, K, kind);
    return B200Q_EINVAL;
  }
  const bool nv = kind == B200Q_KIND_NVF4;
  if (M <= 16) return nv ? launch_decode_t<true, 16>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, static_w, stream)
                         : launch_decode_t<false, 16>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, static_w, stream);
  if (M <= 32) return nv ? launch_decode_t<true, 32>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, static_w, stream)
                         : launch_decode_t<false, 32>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, static_w, stream);
  return nv ? launch_decode_t<true, 64>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, static_w, stream)
            : launch_decode_t<false, 64>(A, B, SFA, SFB, alpha, D, M, N, K, ldd, static_w, stream);
}

}  // namespace b200q

extern "C" int b200q_debug_read_decode_trace(unsigned long long* out, int n) {
  if (n > 16) n = 16;
  B200Q_CUDA(cudaDeviceSynchronize());
  B200Q_CUDA(cudaMemcpyFromSymbol(out, b200q::g_decode_trace, sizeof(unsigned long long) * n));
  return 0;
}
