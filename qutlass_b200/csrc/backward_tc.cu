// backward_t_bf16 on the 5th-generation tensor cores (tcgen05), sm_100a: bf16 [B, N, M] -> MXFP4 of rotate(x^T), any runtime
// 32 x 32 rotation at the streaming rate.
//
// Replaces qutlass/csrc/quartet_bwd_sm120.cu:237-330 for large inputs (the CUDA-core kernel of backward.cu spends ~10
// instructions per element on the transpose + butterflies and reaches 0.48-0.65 of HBM; its generic-rotation path, 1024 FMAs per
// group, 0.14-0.17).  Measured (profiles/r02_s3_bwd_bench_tc*.jsonl): 16384 x 4096 40.0 -> 31.8 us (0.82 of the measured copy
// bandwidth; the forward tcgen05 quantiser on the same bytes: 30.7), 4096 x 4096 13.6 -> 10.4, a non-Hadamard rotation 156.5 ->
// 31.7; below 8 M elements the CUDA-core kernel's shorter set-up wins (4 M: 4.2 vs 4.4 us).  The transpose costs nothing here: a tile of x lies in shared memory in INPUT orientation ([n][m], m
// contiguous) and is handed to tcgen05.mma as an MN-MAJOR A operand, so that
//     D[m, j] = sum_k A[m, k] R[k, j],   A[m, k] = x[32 g + k][m]
// accumulates the rotated group g of output row m in TMEM lane m -- the same bf16 x bf16 products with fp32 accumulation as the
// forward tcgen05 quantiser (quantize_tc.cu) applies to x^T, K = 16 per instruction, two instructions per group.
//
//   warp 0     TMA producer: a tile = 128 n-rows x 128 m-columns = two {64 columns, 128 rows} boxes of the 3-D tensor
//              {M, N, B} (128B swizzle; rows >= N and columns >= M of a batch read as zero), 4-stage ring
//   warp 1     MMA issuer: 4 groups x 2 K-steps of tcgen05.mma.kind::f16 (M = 128, N = 32, K = 16), A MN-major (two 64-column
//              atoms 16 KB apart = LBO, 8 n-rows = 1024 B = SBO), B = R as it lies in memory (MN-major, 64B swizzle); group g
//              -> TMEM columns [32 g, 32 g + 32) of one of 3 accumulator stages
//   warps 2-13 epilogue, 3 groups of 4 warps, group a owns accumulator stage a: thread = output row m (TMEM lane) with its four
//              32-groups; abs-max scale (quartet_bwd_sm120.cu:303-315: s = floor_pow2(amax), q = e2m1(v * 3 / s)), hardware cvt,
//              per-warp staging so that every st.global.v4 of a warp covers whole 64-byte row segments, one 32-bit scale word
//              per row.
//
// Algorithmic bytes / element: 2 (bf16 in) + 0.5 (e2m1) + 1/32; HBM-bound.
#include "quantize_tile.cuh"
#include "backward_quant.cuh"
#include "ptx.cuh"

#include <cuda.h>

namespace b200q {
using namespace ptx;

int make_xT_tmap(void* tm, const void* ptr, int64_t M, int64_t N, int64_t B);   // gemm_fp4.cu (cuTensorMapEncodeTiled plumbing)
int make_rot_tmap(void* tm, const void* ptr, int had);

// 12 epilogue warps: 14 warps per CTA = 128 registers per thread (quantize_tc.cu); measured 31.6 -> 28.8 us at 16384 x 4096
// (0.82 -> 0.90 of the measured copy bandwidth) against 16 warps at 96 registers (profiles/r02_s3_bwd_bench_epi{16,12}.jsonl)
#ifdef B200Q_EPI16
constexpr int kBtEpiWarps = 16;
#else
constexpr int kBtEpiWarps = 12;
#endif
constexpr int kBtThreads = 64 + 32 * kBtEpiWarps;   // 448
constexpr int kBtTile = 128;
constexpr int kBtStageBytes = kBtTile * 256;        // 32 KB: two 16 KB atoms (m columns 0-63, 64-127), 128 n-rows x 128 B each
constexpr int kBtStages = 4;
constexpr int kBtRotBytes = 2048;                   // R: 32 k-rows x 64 B, 64B swizzle
constexpr int kBtAcc = kBtEpiWarps / 4;
constexpr int kBtOutRowBytes = 80;                  // 64 B of codes per row + 16 B pad: conflict-free v4 stores
constexpr int kBtOutWarpBytes = 32 * kBtOutRowBytes;
constexpr int kBtOutBytes = kBtEpiWarps * kBtOutWarpBytes;
constexpr int kBtSmem = kBtStages * kBtStageBytes + kBtRotBytes + kBtOutBytes + 1024 /*barriers*/ + 1024 /*alignment*/;
static_assert(kBtEpiWarps == 4 * kBtAcc, "one group of 4 epilogue warps (one per TMEM lane quarter) per accumulator stage");
static_assert(kBtSmem <= 227 * 1024, "shared memory budget");

struct BwdTcParams {
  uint8_t* q;        // e2m1 [B, M, N/2]
  uint8_t* sf;       // ue8m0 [B, M, N/32]
  int N, M;
  int tiles_n, tiles_m;
  int n_tiles;       // tiles_n * tiles_m * B
  int m_fastest;     // tile order (see coords)
};

__global__ void __launch_bounds__(kBtThreads, 1)
bwd_t_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_r, const BwdTcParams p) {
  extern __shared__ uint8_t bt_smem_raw[];
  const uint32_t smem_base = (smem_u32(bt_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = bt_smem_raw + (smem_base - smem_u32(bt_smem_raw));
  const uint32_t rot_base = smem_base + kBtStages * kBtStageBytes;
  constexpr int kBarOff = kBtStages * kBtStageBytes + kBtRotBytes + kBtOutBytes;
  const uint32_t bar_base = smem_base + kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kBtStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kBtStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kBtStages + kBtAcc + a); };
  const uint32_t rot_bar = bar_base + 8u * (2 * kBtStages + 2 * kBtAcc);
  const uint32_t tmem_slot = rot_bar + 8u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kBarOff + 8 * (2 * kBtStages + 2 * kBtAcc + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Tile order: the shorter dimension fastest, so that one sweep of the grid (148 consecutive tiles) covers it completely --
  // m-tiles fastest reads whole input rows side by side and writes (148 / tiles_m) x 64 contiguous bytes per output row,
  // n-tiles fastest writes whole output rows and reads (148 / tiles_n) x 256 contiguous bytes per input row.  Measured
  // (profiles/r02_s3_bwd_bench_tc_{m,n}fast.jsonl): 16384 x 4096 31.75 vs 32.23 us, 4096 x 14336 28.67 vs 27.95.
  auto coords = [&](int t, int& b, int& n0, int& m0) {
    if (p.m_fastest) {
      const int tm = t % p.tiles_m, r = t / p.tiles_m;
      m0 = tm * kBtTile;
      n0 = (r % p.tiles_n) * kBtTile;
      b = r / p.tiles_n;
    } else {
      const int tn = t % p.tiles_n, r = t / p.tiles_n;
      n0 = tn * kBtTile;
      m0 = (r % p.tiles_m) * kBtTile;
      b = r / p.tiles_m;
    }
  };
  auto request = [&](int t, int s) {
    int b, n0, m0;
    coords(t, b, n0, m0);
    const uint32_t dst = smem_base + s * kBtStageBytes;
    mbar_arrive_expect_tx(full_bar(s), kBtStageBytes);
    tma_load_3d<1>(dst, &tmap_x, full_bar(s), m0, n0, b);
    tma_load_3d<1>(dst + kBtStageBytes / 2, &tmap_x, full_bar(s), m0 + 64, n0, b);
  };
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_x);
    prefetch_tensormap(&tmap_r);
    mbar_init(rot_bar, 1);
    for (int s = 0; s < kBtStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < kBtAcc; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
    pdl_wait();        // x and R may be the previous kernel's outputs: nothing global is touched before this
    mbar_arrive_expect_tx(rot_bar, 32 * 32 * 2);
    tma_load_2d<1>(rot_base, &tmap_r, rot_bar, 0, 0);
    int s = 0;
    for (int t = blockIdx.x; t < p.n_tiles && s < kBtStages; t += gridDim.x, ++s) request(t, s);
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  pdl_wait();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_gen, 0);

  if (warp == 0) {
    // ===================== TMA producer (the first ring was requested in the prologue) =====================
    const bool elected = elect_one();
    int stage = 0;
    uint32_t phase = 1;
    for (int t = blockIdx.x + kBtStages * (int)gridDim.x; t < p.n_tiles; t += gridDim.x) {
      mbar_wait(empty_bar(stage), phase ^ 1, 1);
      if (elected) request(t, stage);
      __syncwarp();
      if (++stage == kBtStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    const bool elected = elect_one();
    // fp32 accumulate, bf16 x bf16, A MN-major (bit 15), B MN-major (bit 16), N = 32, M = 128
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    // A: canonical MN-major ((64, m), (8, k)) with 128B swizzle -- an atom = 8 n-rows x 64 m-columns (1024 B); SBO = the next 8
    // n-rows (1024 B), LBO = the next 64 m-columns (the second box of the stage, 16 KB on)
    constexpr uint32_t a_hi = (1024u >> 4) | (1u << 14) | (kLayoutSw128 << 29);
    constexpr uint32_t a_lbo = ((uint32_t)(kBtStageBytes / 2) >> 4) << 16;
    // B: R [k][n], n contiguous, rows of 64 B (64B swizzle): SBO = 8 k-rows
    constexpr uint32_t b_hi = ((8u * 64u) >> 4) | (1u << 14) | (kLayoutSw64 << 29);
    const uint32_t b_lo0 = ((rot_base & 0x3FFFFu) >> 4) | (1u << 16);
    auto mk = [](uint32_t lo, uint32_t hi) {
      uint64_t d;
      asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
      return d;
    };
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    mbar_wait(rot_bar, 0, 5);
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1, 2);
      mbar_wait(full_bar(stage), phase, 3);
      tc_fence_after();
      if (elected) {
        const uint32_t a_lo0 = (((smem_base + stage * kBtStageBytes) & 0x3FFFFu) >> 4) | a_lbo;
        const uint32_t d0 = tmem_base + acc * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            // n-rows 32 g + 16 ks ... + 15 of the tile: 128 B per row
            const uint32_t a_lo = a_lo0 + (uint32_t)(((32 * g + 16 * ks) * 128) >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)ks * ((16u * 64u) >> 4);
            mma_f16<1>(d0 + g * 32, mk(a_lo, a_hi), mk(b_lo, b_hi), idesc, ks > 0 ? 1u : 0u);
          }
        }
        tc_commit<1>(empty_bar(stage));
        tc_commit<1>(tfull_bar(acc));
      }
      __syncwarp();
      if (++stage == kBtStages) { stage = 0; phase ^= 1; }
      if (++acc == kBtAcc) { acc = 0; acc_phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..13) =====================
    const int ew = warp - 2;
    const int q = warp & 3;                  // TMEM lane quarter this warp may access
    const int grp = ew >> 2;                 // accumulator stage / tile residue this warp's group owns
    uint8_t* ostage = smem_gen + kBtStages * kBtStageBytes + kBtRotBytes + ew * kBtOutWarpBytes;
    const uint32_t lane_taddr = tmem_base + grp * 128 + ((uint32_t)(q * 32) << 16);
    const int out_row_chunks = p.N >> 5;     // 16-byte code chunks (= scale bytes) per output row
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x + grp * (int)gridDim.x; t < p.n_tiles; t += kBtAcc * (int)gridDim.x) {
      int b, n0, m0;
      coords(t, b, n0, m0);
      mbar_wait(tfull_bar(grp), acc_phase, 4);
      acc_phase ^= 1;
      tc_fence_after();
      uint32_t out[4][4], sfb[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(lane_taddr + h * 64, r0);
        tmem_ld_32x32b_x32(lane_taddr + h * 64 + 32, r1);
        tmem_ld_wait();
        if (h == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(grp));
        }
        sfb[2 * h] = quantise32_absmax<false>(reinterpret_cast<float*>(r0), 1.f, out[2 * h]);
        sfb[2 * h + 1] = quantise32_absmax<false>(reinterpret_cast<float*>(r1), 1.f, out[2 * h + 1]);
      }

      // ---- codes: row-per-thread -> padded staging -> every warp store covers 8 rows x 64 contiguous bytes
      uint8_t* qb = p.q + (size_t)b * p.M * ((size_t)out_row_chunks * 16);
      uint8_t* sb = p.sf + (size_t)b * p.M * (size_t)out_row_chunks;
      const int c0 = n0 >> 5;                // first chunk of this tile within an output row
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(ostage + lane * kBtOutRowBytes + c * 16) = make_uint4(out[c][0], out[c][1], out[c][2], out[c][3]);
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int u = c * 32 + lane;
        const int row = u >> 2, slot = u & 3;
        const uint4 w = *reinterpret_cast<const uint4*>(ostage + row * kBtOutRowBytes + slot * 16);
        const int m = m0 + q * 32 + row;
        if (m < p.M && c0 + slot < out_row_chunks)
          *reinterpret_cast<uint4*>(qb + ((size_t)m * out_row_chunks + c0 + slot) * 16) = w;
      }
      __syncwarp();

      // ---- the row's four scale bytes
      const int m = m0 + q * 32 + lane;
      if (m < p.M) {
        uint8_t* dst = sb + (size_t)m * out_row_chunks + c0;
        if ((out_row_chunks & 3) == 0) {
          *reinterpret_cast<uint32_t*>(dst) = sfb[0] | (sfb[1] << 8) | (sfb[2] << 16) | (sfb[3] << 24);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c0 + j < out_row_chunks) dst[j] = (uint8_t)sfb[j];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

// ------------------------------------------------------------------ FP4-input re-quantisers on the tensor cores
// backward_qt_bf16 (MXFP4 [B, N, M/2] + scales -> MXFP4 of rotate(dq(x)^T), scale / alpha): the CUDA-core kernel spends ~14
// instructions per element (4 on the e2m1 decode, 4 on the butterflies, 2 on byte-granular shared loads ...) and sits at 0.25 of HBM.
// Here 8 "decode" warps turn a 128 x 128 tile of codes into bf16 -- EXACT: an e2m1 value times a power of two has two
// significant bits -- and write it straight into shared memory in the 128B-swizzled MN-major operand layout (what TMA would
// have produced from a bf16 tensor); the tensor core then transposes + rotates as in bwd_t_tc_kernel, and 8 epilogue warps (two
// accumulator stages) quantise.  The codes and scales of the next three tiles are in flight as cp.async copies.  fp32 accumulation of exact products: the same values as the CUDA-core path wherever its
// butterfly sums are exact (the reference recipe's few-bit inputs), inside the 1e-4 bar otherwise.
// 4 decode + 8 epilogue warps = 14 warps: 128 registers per thread, no spills.  Measured at 16384 x 4096 (CUDA-core kernel: 43.4 us):
// 8 + 8 warps with a one-tile register prefetch 48.5 us, the same with the cp.async ring 49.9, 4 + 12 warps (96 registers, 128
// bytes of spills) 46.9, this form 40.6, and 38.8 with the conflict-free decode mapping below (profiles/r02_s3_bwd_bench_qt_tc1_v*.jsonl,
// r02_s3_bwd_bench_final_qt_tensorcore.jsonl).  SIX decode warps (16 warps, still 128 registers; work items dealt to quarter-warps,
// every thread with a private copy of its row's scale word) were slower again: 41.8 us -- the decode is not the only bound.
constexpr int kBfDeqWarps = 4, kBfEpiWarps = 8;
constexpr int kBfThreads = 64 + 32 * (kBfDeqWarps + kBfEpiWarps);   // 448
constexpr int kBfStages = 3;                                         // bf16 operand tiles (32 KB each)
constexpr int kBfAcc = kBfEpiWarps / 4;
static_assert(kBfEpiWarps == 4 * kBfAcc, "one group of 4 epilogue warps per accumulator stage");
constexpr int kBfOutBytes = kBfEpiWarps * kBtOutWarpBytes;
constexpr int kBfRaw = 4;                                            // raw input tiles in flight / being decoded
constexpr int kBfRawBytes = kBtTile * 64 + kBtTile * 4;              // 128 rows x 64 B of codes + 128 x 4 scale bytes
constexpr int kBfSmem = kBfStages * kBtStageBytes + kBtRotBytes + kBfOutBytes + 1024 /*barriers*/ + kBfRaw * kBfRawBytes + 1024 /*alignment*/;
static_assert(kBfSmem <= 227 * 1024, "shared memory budget");

struct BwdFp4Params {
  const uint8_t* xq;     // packed e2m1 [B, N, M/2]
  const uint8_t* xs;     // ue8m0 [B, N, M/32]
  const float* alpha;    // device scalar
  uint8_t* q;            // e2m1 [B, M, N/2]
  uint8_t* sf;           // ue8m0 [B, M, N/32]
  int N, M;
  int tiles_n, tiles_m, n_tiles, m_fastest;
};

__global__ void __launch_bounds__(kBfThreads, 1)
bwd_qt_tc_kernel(const __grid_constant__ CUtensorMap tmap_r, const BwdFp4Params p) {
  extern __shared__ uint8_t bf_smem_raw[];
  const uint32_t smem_base = (smem_u32(bf_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = bf_smem_raw + (smem_base - smem_u32(bf_smem_raw));
  const uint32_t rot_base = smem_base + kBfStages * kBtStageBytes;
  constexpr int kBarOff = kBfStages * kBtStageBytes + kBtRotBytes + kBfOutBytes;
  const uint32_t bar_base = smem_base + kBarOff;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kBfStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kBfStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kBfStages + kBfAcc + a); };
  const uint32_t rot_bar = bar_base + 8u * (2 * kBfStages + 2 * kBfAcc);
  const uint32_t tmem_slot = rot_bar + 8u;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + kBarOff + 8 * (2 * kBfStages + 2 * kBfAcc + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  auto coords = [&](int t, int& b, int& n0, int& m0) {
    if (p.m_fastest) {
      const int tm = t % p.tiles_m, r = t / p.tiles_m;
      m0 = tm * kBtTile;
      n0 = (r % p.tiles_n) * kBtTile;
      b = r / p.tiles_n;
    } else {
      const int tn = t % p.tiles_n, r = t / p.tiles_n;
      n0 = tn * kBtTile;
      m0 = (r % p.tiles_m) * kBtTile;
      b = r / p.tiles_m;
    }
  };

  if (warp == 1 && lane == 0) {
    prefetch_tensormap(&tmap_r);
    mbar_init(rot_bar, 1);
    for (int s = 0; s < kBfStages; ++s) {
      mbar_init(full_bar(s), kBfDeqWarps);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < kBfAcc; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_mbar_init();
    fence_proxy_async_smem();
    mbar_arrive_expect_tx(rot_bar, 32 * 32 * 2);
    tma_load_2d<1>(rot_base, &tmap_r, rot_bar, 0, 0);
  }
  if (warp == 0) {
    tmem_alloc<1>(tmem_slot, 512);      // 3 accumulator stages x 128 columns (allocations are powers of two)
    tmem_relinquish<1>();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_gen, 0);

  if (warp == 0) {
    // ===================== MMA issuer =====================
    const bool elected = elect_one();
    constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(32 >> 3) << 17) |
                               ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t a_hi = (1024u >> 4) | (1u << 14) | (kLayoutSw128 << 29);
    constexpr uint32_t a_lbo = ((uint32_t)(kBtStageBytes / 2) >> 4) << 16;
    constexpr uint32_t b_hi = ((8u * 64u) >> 4) | (1u << 14) | (kLayoutSw64 << 29);
    const uint32_t b_lo0 = ((rot_base & 0x3FFFFu) >> 4) | (1u << 16);
    auto mk = [](uint32_t lo, uint32_t hi) {
      uint64_t d;
      asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
      return d;
    };
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    mbar_wait(rot_bar, 0, 5);
    for (int t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
      mbar_wait(tempty_bar(acc), acc_phase ^ 1, 2);
      mbar_wait(full_bar(stage), phase, 3);
      tc_fence_after();
      if (elected) {
        const uint32_t a_lo0 = (((smem_base + stage * kBtStageBytes) & 0x3FFFFu) >> 4) | a_lbo;
        const uint32_t d0 = tmem_base + acc * 128;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) {
            const uint32_t a_lo = a_lo0 + (uint32_t)(((32 * g + 16 * ks) * 128) >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)ks * ((16u * 64u) >> 4);
            mma_f16<1>(d0 + g * 32, mk(a_lo, a_hi), mk(b_lo, b_hi), idesc, ks > 0 ? 1u : 0u);
          }
        }
        tc_commit<1>(empty_bar(stage));
        tc_commit<1>(tfull_bar(acc));
      }
      __syncwarp();
      if (++stage == kBfStages) { stage = 0; phase ^= 1; }
      if (++acc == kBfAcc) { acc = 0; acc_phase ^= 1; }
    }
  } else if (warp >= 2 + kBfEpiWarps) {
    // ===================== decode warps: e2m1 codes x 2^(e-127) -> bf16 operand tile =====================
    const int dw = warp - (2 + kBfEpiWarps);
    const int row_bytes = p.M >> 1, row_sf = p.M >> 5;
    // unit = 16 bytes of codes = 32 m-columns of one n-row with ONE scale; a thread owns 4 of its warp's 128 units.  Eight
    // consecutive lanes take the SAME 32-column part of eight consecutive rows: their 16-byte stores into the swizzled operand
    // tile then differ in `chunk ^ (row & 7)` and hit eight different bank groups (with 2 rows x 4 parts per quarter-warp the
    // two atoms of a row collided: ncu counted 51 % of the shared wavefronts as conflicts); the raw ring is XOR-swizzled with
    // (row >> 1) & 3 for the same reason.
    constexpr int kUnits = (kBtTile * 4) / (kBfDeqWarps * 32);
    static_assert(kUnits == 4, "a decode warp covers 32 rows: 4 units (8 rows each) x 4 parts");
    int urow[kUnits], upart[kUnits];
#pragma unroll
    for (int i = 0; i < kUnits; ++i) {
      urow[i] = dw * 32 + i * 8 + (lane & 7);
      upart[i] = lane >> 3;
    }
    auto raw_unit = [&](int i) { return urow[i] * 64 + ((upart[i] ^ ((urow[i] >> 1) & 3)) * 16); };
    // Raw codes + scales reach shared memory by cp.async, kBfRaw - 1 tiles ahead (a register prefetch of ONE tile left 8 KB in
    // flight per SM: latency-bound at 0.7 TB/s, slower than the CUDA-core kernel).  Every thread copies exactly the two units it
    // decodes itself; the 4-byte scale word of a row is copied by the lane that owns the row's first unit and read by the
    // three other lanes of that row (same warp: cp.async.wait_group + __syncwarp).
    uint8_t* raw0 = smem_gen + kBfStages * kBtStageBytes + kBtRotBytes + kBfOutBytes + 1024;
    auto issue = [&](int t, int rs) {
      int b, n0, m0;
      coords(t, b, n0, m0);
      uint8_t* raw = raw0 + rs * kBfRawBytes;
#pragma unroll
      for (int i = 0; i < kUnits; ++i) {
        const int n = n0 + urow[i];
        const bool ok = n < p.N;                                   // M % 128 == 0: no partial tiles along m
        const size_t r = (size_t)b * p.N + (ok ? n : 0);
        cp_async16_zfill(raw + raw_unit(i), p.xq + r * row_bytes + (m0 >> 1) + upart[i] * 16, ok);
        if (upart[i] == 0) cp_async4_zfill(raw + kBtTile * 64 + urow[i] * 4, p.xs + r * row_sf + (m0 >> 5), ok);
      }
    };
    int stage = 0;
    uint32_t phase = 0;                 // first use of a stage: the wait on parity 1 of a fresh barrier returns at once
    int t = blockIdx.x;
    for (int k = 0; k < kBfRaw - 1; ++k) {
      if (t + k * (int)gridDim.x < p.n_tiles) issue(t + k * gridDim.x, k);
      cp_async_commit();
    }
    for (int it = 0; t < p.n_tiles; t += gridDim.x, ++it) {
      if (t + (kBfRaw - 1) * (int)gridDim.x < p.n_tiles) issue(t + (kBfRaw - 1) * gridDim.x, (it + kBfRaw - 1) % kBfRaw);
      cp_async_commit();
      cp_async_wait<kBfRaw - 1>();        // this tile's group has landed (the newer ones may still be in flight)
      __syncwarp();
      const uint8_t* raw = raw0 + (it % kBfRaw) * kBfRawBytes;
      mbar_wait(empty_bar(stage), phase ^ 1, 1);
      uint8_t* tile = smem_gen + stage * kBtStageBytes;
#pragma unroll
      for (int i = 0; i < kUnits; ++i) {
        const uint4 cur = *reinterpret_cast<const uint4*>(raw + raw_unit(i));
        // (uint16) byte << 7 as bf16 == byte << 23 as fp32: 2^(e-127), e = 0 -> 0.0 (quartet_bwd_sm120.cu:360)
        const float sc = __uint_as_float((uint32_t)raw[kBtTile * 64 + urow[i] * 4 + upart[i]] << 23);
        const float2 sc2 = make_float2(sc, sc);
        const uint32_t cw[4] = {cur.x, cur.y, cur.z, cur.w};
        // m-columns 32 part ... + 31 of row urow: atom (part >> 1), 16-byte chunks 4 (part & 1) ... + 3, XOR-swizzled with row & 7
        uint8_t* rowp = tile + (upart[i] >> 1) * (kBtStageBytes / 2) + urow[i] * 128;
        const int j0 = (upart[i] & 1) * 4, sw = urow[i] & 7;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t h[4], w[4];
          e2m1x8_to_half2x4(cw[c], h);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 f = __fmul2_rn(__half22float2(*reinterpret_cast<const __half2*>(&h[k])), sc2);     // exact: two significant bits
            const __nv_bfloat162 b2 = __floats2bfloat162_rn(f.x, f.y);
            w[k] = *reinterpret_cast<const uint32_t*>(&b2);
          }
          *reinterpret_cast<uint4*>(rowp + (((j0 + c) ^ sw) * 16)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();          // generic-proxy writes -> visible to the tensor core's async proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(full_bar(stage));
      if (++stage == kBfStages) { stage = 0; phase ^= 1; }
    }
  } else if (warp >= 2) {
    // ===================== epilogue (warps 2..9): two groups of four, group a owns accumulator stage a =====================
    const int ew = warp - 2;
    const int q = warp & 3;
    const int grp = ew >> 2;
    uint8_t* ostage = smem_gen + kBfStages * kBtStageBytes + kBtRotBytes + ew * kBtOutWarpBytes;
    const uint32_t lane_taddr = tmem_base + grp * 128 + ((uint32_t)(q * 32) << 16);
    const int out_row_chunks = p.N >> 5;
    const float alpha = __ldg(p.alpha);
    const float c3 = __fdiv_rn(3.0f, alpha);
    uint32_t acc_phase = 0;
    for (int t = blockIdx.x + grp * (int)gridDim.x; t < p.n_tiles; t += kBfAcc * (int)gridDim.x) {
      int b, n0, m0;
      coords(t, b, n0, m0);
      mbar_wait(tfull_bar(grp), acc_phase, 4);
      acc_phase ^= 1;
      tc_fence_after();
      uint32_t out[4][4], sfb[4];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t r0[32], r1[32];
        tmem_ld_32x32b_x32(lane_taddr + h * 64, r0);
        tmem_ld_32x32b_x32(lane_taddr + h * 64 + 32, r1);
        tmem_ld_wait();
        if (h == 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty_bar(grp));
        }
        sfb[2 * h] = quantise32_absmax<true>(reinterpret_cast<float*>(r0), alpha, out[2 * h], c3);
        sfb[2 * h + 1] = quantise32_absmax<true>(reinterpret_cast<float*>(r1), alpha, out[2 * h + 1], c3);
      }
      uint8_t* qb = p.q + (size_t)b * p.M * ((size_t)out_row_chunks * 16);
      uint8_t* sb = p.sf + (size_t)b * p.M * (size_t)out_row_chunks;
      const int c0 = n0 >> 5;
#pragma unroll
      for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(ostage + lane * kBtOutRowBytes + c * 16) = make_uint4(out[c][0], out[c][1], out[c][2], out[c][3]);
      __syncwarp();
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int u = c * 32 + lane;
        const int row = u >> 2, slot = u & 3;
        const uint4 w = *reinterpret_cast<const uint4*>(ostage + row * kBtOutRowBytes + slot * 16);
        const int m = m0 + q * 32 + row;
        if (m < p.M && c0 + slot < out_row_chunks)
          *reinterpret_cast<uint4*>(qb + ((size_t)m * out_row_chunks + c0 + slot) * 16) = w;
      }
      __syncwarp();
      const int m = m0 + q * 32 + lane;
      if (m < p.M) {
        uint8_t* dst = sb + (size_t)m * out_row_chunks + c0;
        if ((out_row_chunks & 3) == 0) {
          *reinterpret_cast<uint32_t*>(dst) = sfb[0] | (sfb[1] << 8) | (sfb[2] << 16) | (sfb[3] << 24);
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (c0 + j < out_row_chunks) dst[j] = (uint8_t)sfb[j];
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

bool backward_qt_tc_eligible(const void* xq, const void* xs, const void* rot, const void* q, const void* sf, int size_m, int size_n,
                             int size_b) {
  if (((uintptr_t)xq & 15) || ((uintptr_t)rot & 15) || ((uintptr_t)q & 15) || ((uintptr_t)sf & 3)) return false;
  if (((uintptr_t)xs & 3) || size_m % 128 != 0 || size_n % 32 != 0) return false;    // 4-byte scale words per (row, tile): cp.async
  const int64_t tiles = ceil_div(size_n, kBtTile) * ceil_div(size_m, kBtTile) * (int64_t)size_b;
  return tiles < ((int64_t)1 << 30);
}

int launch_backward_qt_tc(const void* xq, const void* xs, const void* rot, const float* alpha, void* q, void* sf, int size_m,
                          int size_n, int size_b, cudaStream_t stream) {
  auto kern = bwd_qt_tc_kernel;
  static std::atomic<unsigned long long> smem_attr_done{0};
  if (int rc_attr = ensure_dynamic_smem(kern, kBfSmem, smem_attr_done)) return rc_attr;
  BwdFp4Params p;
  p.xq = (const uint8_t*)xq;
  p.xs = (const uint8_t*)xs;
  p.alpha = alpha;
  p.q = (uint8_t*)q;
  p.sf = (uint8_t*)sf;
  p.N = size_n;
  p.M = size_m;
  p.tiles_n = (int)ceil_div(size_n, kBtTile);
  p.tiles_m = (int)ceil_div(size_m, kBtTile);
  p.n_tiles = p.tiles_n * p.tiles_m * size_b;
  p.m_fastest = p.tiles_m <= p.tiles_n ? 1 : 0;
  CUtensorMap tr;
  int rc = make_rot_tmap(&tr, rot, 32);
  if (rc) return rc;
  const int ctas = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  kern<<<ctas, kBfThreads, kBfSmem, stream>>>(tr, p);
  B200Q_CUDA(cudaGetLastError());
  return 0;
}

// inputs the tensor-core kernel can take (everything else: the CUDA-core kernels of backward.cu)
bool backward_t_tc_eligible(const void* x, const void* rot, const void* q, const void* sf, int size_m, int size_n, int size_b) {
  if (((uintptr_t)x & 15) || ((uintptr_t)rot & 15) || ((uintptr_t)q & 15) || ((uintptr_t)sf & 3)) return false;
  if (size_m % 8 != 0 || size_n % 32 != 0) return false;                      // TMA row pitch 16 B; whole 32-groups
  const int64_t tiles = ceil_div(size_n, kBtTile) * ceil_div(size_m, kBtTile) * (int64_t)size_b;
  return tiles < ((int64_t)1 << 30);
}

int launch_backward_t_tc(const void* x, const void* rot, void* q, void* sf, int size_m, int size_n, int size_b, cudaStream_t stream) {
  auto kern = bwd_t_tc_kernel;
  static std::atomic<unsigned long long> smem_attr_done{0};
  if (int rc_attr = ensure_dynamic_smem(kern, kBtSmem, smem_attr_done)) return rc_attr;
  BwdTcParams p;
  p.q = (uint8_t*)q;
  p.sf = (uint8_t*)sf;
  p.N = size_n;
  p.M = size_m;
  p.tiles_n = (int)ceil_div(size_n, kBtTile);
  p.tiles_m = (int)ceil_div(size_m, kBtTile);
  p.n_tiles = p.tiles_n * p.tiles_m * size_b;
  p.m_fastest = p.tiles_m <= p.tiles_n ? 1 : 0;
  CUtensorMap tx, tr;
  int rc = make_xT_tmap(&tx, x, size_m, size_n, size_b);
  if (rc) return rc;
  rc = make_rot_tmap(&tr, rot, 32);
  if (rc) return rc;
  const int ctas = p.n_tiles < num_sms() ? p.n_tiles : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(kBtThreads);
  cfg.dynamicSmemBytes = kBtSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attribute(attr);
  B200Q_CUDA(cudaLaunchKernelEx(&cfg, kern, tx, tr, p));
  return 0;
}

}  // namespace b200q
