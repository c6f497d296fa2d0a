// Fused rotate + microscaled-FP4 quantise with the rotation on the 5th-generation tensor cores (tcgen05), sm_100a.
//
// Same contract as quantize.cu (replaces qutlass/csrc/fused_quantize_{mx,nv,mx_mask}.cu +
// cutlass_extensions/epilogue/threadblock/epilogue_quant.h: a bf16 GEMM x_group(1xH) @ R(HxH) with fp32 accumulation and
// a quantising epilogue), but the CUDA cores only run the epilogue: the butterfly kernel spends ~14 instructions per
// element on the rotation and is issue-bound at 0.45-0.6 of HBM; here the rotation costs nothing and ANY runtime R is
// handled at full speed -- exactly the reference's arithmetic (bf16 x bf16 products, fp32 accumulate).
//
// The flat activation tensor is viewed as [numel / 128, 128] bf16.  One tile = 128 such rows (32 KB):
//   warp 0     TMA producer: two 128-row x 64-element boxes (128B swizzle) per tile into a 5-stage ring
//   warp 1     MMA issuer: 128 / H groups x H / 16 K-steps of tcgen05.mma.kind::f16 (M = 128, N = H, K = 16): group g
//              multiplies columns [gH, gH + H) of the tile with R (TMA-loaded once as it lies in memory = an MN-major B operand) into TMEM
//              columns [gH, gH + H) of one of 3 accumulator stages -- no zero padding, so a NaN / inf stays in its group
//   warps 2-13 epilogue, 3 groups of 4 warps; group g owns accumulator stage g, i.e. every 3rd tile of the CTA, so three
//              tiles are in flight in the CUDA cores.  thread = one 128-element row (TMEM lane): 4 chunks of 32 columns,
//              two at a time (tcgen05.ld x32 twice -> two independent scale / e2m1 chains), accumulator released after
//              the last load.  A thread's 64 bytes of codes go through a padded per-warp staging buffer so that every
//              st.global.v4 of the warp covers 512 contiguous bytes; the 4 (NV: 8) scale bytes of a row are ONE 32-bit
//              (64-bit) store to the row-major buffer and one / two 32-bit stores to the blocked buffer.
//              (First version: warp = lane quarter x 32-column chunk, every warp visiting every tile; its 16-byte codes
//              went out at a 64-byte lane stride and its scale bytes one by one -- ncu: 46 % of the stall samples sat
//              behind those stores, 4.1 TB/s.)
//
// Algorithmic bytes / element: 2 (bf16 in) + 0.5 (e2m1) + 1/32 (+1/32 blocked)  [MX]; HBM-bound.
#include "quantize_tile.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace b200q {
using namespace ptx;

int make_x128_tmap(void* tm, const void* ptr, int64_t rows);   // gemm_fp4.cu (cuTensorMapEncodeTiled plumbing)
int make_rot_tmap(void* tm, const void* ptr, int had);

// 12 epilogue warps (three groups of four): 14 warps per CTA leave every thread 104-128 registers (four warps per scheduler), so
// the epilogue neither spills nor serialises; with 16 (18 warps, five on one scheduler) ptxas caps the kernel at 96 registers
// and spills 32-64 bytes per thread.  Measured (profiles/r02_s3_quant_sweep_epi{16,12}.jsonl): quest H = 128 10.99 -> 10.56 us at
// 4096 x 4096, 30.4 -> 29.8 at 16384 rows; abs_max unchanged.  -DB200Q_EPI16 builds the former form.
#ifdef B200Q_EPI16
constexpr int kTcEpiWarps = 16;
#else
constexpr int kTcEpiWarps = 12;
#endif
constexpr int kTcThreads = 64 + 32 * kTcEpiWarps;   // 448
constexpr int kTcTileRows = 128;
constexpr int kTcStageBytes = kTcTileRows * 256;    // 32 KB: two 16 KB swizzle atoms columns (K halves)
constexpr int kTcStages = 4;
constexpr int kTcRotBytes = 128 * 256;              // R^T for H = 128: two atoms of 128 rows x 128 B
constexpr int kTcAcc = kTcEpiWarps / 4;             // accumulator stages of 128 TMEM columns
constexpr int kTcOutRowBytes = 80;                  // 64 B of codes per row + 16 B pad: conflict-free v4 stores
constexpr int kTcOutWarpBytes = 32 * kTcOutRowBytes;
constexpr int kTcOutBytes = kTcEpiWarps * kTcOutWarpBytes;
constexpr int kTcBarBytes = 1024;
constexpr int kTcSmem = kTcStages * kTcStageBytes + kTcRotBytes + kTcOutBytes + kTcBarBytes + 1024;   // + alignment slack
static_assert(kTcEpiWarps == 4 * kTcAcc, "one group of 4 epilogue warps (one per TMEM lane quarter) per accumulator stage");
static_assert(kTcSmem <= 227 * 1024, "shared memory budget");

template <bool NV, int METHOD, bool MASK>
__global__ void __launch_bounds__(kTcThreads, 1)
quantize_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_r, const QuantParams p,
                   const int had, const int64_t n_tiles) {
  extern __shared__ uint8_t tc_smem_raw[];
  const uint32_t smem_base = (smem_u32(tc_smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = tc_smem_raw + (smem_base - smem_u32(tc_smem_raw));
  const uint32_t rot_base = smem_base + kTcStages * kTcStageBytes;
  const uint32_t bar_base = rot_base + kTcRotBytes + kTcOutBytes;
  constexpr int kBarOff = kTcStages * kTcStageBytes + kTcRotBytes + kTcOutBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kTcStages + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * kTcStages + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * kTcStages + kTcAcc + a); };
  const uint32_t rot_bar = bar_base + 8u * (2 * kTcStages + 2 * kTcAcc);
  const uint32_t tmem_slot = rot_bar + 8u;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kBarOff + 8 * (2 * kTcStages + 2 * kTcAcc + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // a dependent kernel launched with the PDL attribute (our GEMM) may start its prologue / weight loads now
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_x);
    prefetch_tensormap(&tmap_r);
    mbar_init(rot_bar, 1);
    for (int s = 0; s < kTcStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < kTcAcc; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);
    }
    fence_mbar_init();
    fence_proxy_async_smem();      // the initialised barriers as the TMA unit (async proxy) must see them
    // Launched as a programmatic dependent (launch_tc): everything above overlapped the tail of the previous kernel in the
    // stream; x, R and global_scale may be its outputs and our output buffers may still be read by it, so every access
    // to global memory sits behind griddepcontrol.wait (a no-op for a plain launch).
    pdl_wait();
    // First ring of loads right here, before the TMEM allocation and the CTA-wide barrier: the ring is empty, and the
    // DRAM latency of the first tile is the longest item of the kernel's fixed cost.  R as it lies in memory ([k][n],
    // n contiguous) is an MN-major B operand: rows of min(H, 64) elements, swizzle span = row bytes (32 / 64 / 128 B);
    // H = 128 takes two 64-column boxes, H * 128 B apart.
    mbar_arrive_expect_tx(rot_bar, (uint32_t)(had * had * 2));
    tma_load_2d<1>(rot_base, &tmap_r, rot_bar, 0, 0);
    if (had == 128) tma_load_2d<1>(rot_base + 128 * 128, &tmap_r, rot_bar, 64, 0);
    int s = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles && s < kTcStages; tile += gridDim.x, ++s) {
      const uint32_t dst = smem_base + s * kTcStageBytes;
      mbar_arrive_expect_tx(full_bar(s), kTcStageBytes);
      tma_load_2d<1>(dst, &tmap_x, full_bar(s), 0, (int32_t)(tile * kTcTileRows));
      tma_load_2d<1>(dst + kTcStageBytes / 2, &tmap_x, full_bar(s), 64, (int32_t)(tile * kTcTileRows));
    }
  }
  if (warp == 1) {
    tmem_alloc<1>(tmem_slot, 512);
    tmem_relinquish<1>();
  }
  tc_fence_before();
  pdl_wait();          // every thread: nothing below may touch global memory before the previous kernel has completed
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot_gen, 0);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // (the first kTcStages tiles and R were requested in the prologue above)
    const bool elected = elect_one();
    int stage = 0;
    uint32_t phase = 1;
    for (int64_t tile = blockIdx.x + (int64_t)kTcStages * gridDim.x; tile < n_tiles; tile += gridDim.x) {
      mbar_wait(empty_bar(stage), phase ^ 1, 1);
      if (elected) {
        const uint32_t dst = smem_base + stage * kTcStageBytes;
        mbar_arrive_expect_tx(full_bar(stage), kTcStageBytes);
        tma_load_2d<1>(dst, &tmap_x, full_bar(stage), 0, (int32_t)(tile * kTcTileRows));
        tma_load_2d<1>(dst + kTcStageBytes / 2, &tmap_x, full_bar(stage), 64, (int32_t)(tile * kTcTileRows));
      }
      __syncwarp();
      if (++stage == kTcStages) { stage = 0; phase ^= 1; }
    }
  } else {
    if (warp == 1) {
      // ===================== MMA issuer =====================
      const bool elected = elect_one();
      const int groups = 128 / had, ksteps = had >> 4;
      // fp32 accumulate, bf16 x bf16, B MN-major (bit 16), N = H, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(had >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (kLayoutSw128 << 29);   // A: SBO 1024 B, version 1, 128B swizzle
      // B: canonical MN-major ((8,8,m),(8,k)):((1,8,LBO),(row,SBO)) -- SBO = 8 k-rows, LBO = next 64 columns (H = 128 only)
      const uint32_t b_row = (uint32_t)(had < 64 ? had : 64) * 2u;
      const uint32_t b_layout = had >= 64 ? kLayoutSw128 : (had == 32 ? kLayoutSw64 : kLayoutSw32);
      const uint32_t b_hi = ((8u * b_row) >> 4) | (1u << 14) | (b_layout << 29);
      const uint32_t b_lo0 = ((rot_base & 0x3FFFFu) >> 4) | ((had == 128 ? (128u * 128u) >> 4 : 1u) << 16);
      const uint32_t b_kstep16 = (16u * b_row) >> 4;                                   // 16 k-rows per MMA
      auto mk = [](uint32_t lo, uint32_t hi) {
        uint64_t d;
        asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(hi));
        return d;
      };
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      mbar_wait(rot_bar, 0, 5);
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        mbar_wait(tempty_bar(acc), acc_phase ^ 1, 2);
        mbar_wait(full_bar(stage), phase, 3);
        tc_fence_after();
        if (elected) {
          const uint32_t a_lo0 = (((smem_base + stage * kTcStageBytes) & 0x3FFFFu) >> 4) | (1u << 16);
          const uint32_t d0 = tmem_base + acc * 128;
          for (int g = 0; g < groups; ++g) {
            for (int ks = 0; ks < ksteps; ++ks) {
              const int kg = g * had + ks * 16;
              const uint32_t a_lo = a_lo0 + (uint32_t)(kg >> 6) * (uint32_t)(kTcStageBytes / 2 / 16) + (uint32_t)((kg & 63) >> 4) * 2u;
              const uint32_t b_lo = b_lo0 + (uint32_t)ks * b_kstep16;
              mma_f16<1>(d0 + g * had, mk(a_lo, kDescHi), mk(b_lo, b_hi), idesc, ks > 0 ? 1u : 0u);
            }
          }
          tc_commit<1>(empty_bar(stage));    // stage free once the MMAs have read it
          tc_commit<1>(tfull_bar(acc));      // accumulator complete
        }
        __syncwarp();
        if (++stage == kTcStages) { stage = 0; phase ^= 1; }
        if (++acc == kTcAcc) { acc = 0; acc_phase ^= 1; }
      }
    } else {
      // ===================== epilogue (warps 2..13) =====================
      const int ew = warp - 2;
      const int q = warp & 3;                  // TMEM lane quarter this warp may access
      const int grp = ew >> 2;                 // accumulator stage / tile residue this warp's group owns
      uint8_t* ostage = smem_gen + kTcStages * kTcStageBytes + kTcRotBytes + ew * kTcOutWarpBytes;
      const bool nvq = NV && METHOD == B200Q_METHOD_ABSMAX && p.nv_sm100_codes != 0;
      float gs = 1.f, gs_rcp = 1.f;
      if constexpr (NV) {
        gs = *p.gs;
        gs_rcp = rcp_approx_ftz(gs);
      }
      // all scale bytes of a 128-element row share one 4-byte cell of the blocked layout (NV: two cells, 512 B apart)
      const bool row_cells = NV ? ((p.cols & 7) == 0) : ((p.cols & 3) == 0);
      const uint32_t lane_taddr = tmem_base + grp * 128 + ((uint32_t)(q * 32) << 16);
      uint32_t acc_phase = 0;
      for (int64_t tile = blockIdx.x + (int64_t)grp * gridDim.x; tile < n_tiles; tile += (int64_t)kTcAcc * gridDim.x) {
        mbar_wait(tfull_bar(grp), acc_phase, 4);
        acc_phase ^= 1;
        tc_fence_after();
        uint32_t out[4][4], sfb[4], mk[4];
        if constexpr (!NV) {
          // two chunks per TMEM round trip: two independent scale / e2m1 chains in flight
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
            chunk_quantise<NV, METHOD, MASK>(reinterpret_cast<float*>(r0), gs, gs_rcp, out[2 * h], sfb[2 * h], mk[2 * h], nvq);
            chunk_quantise<NV, METHOD, MASK>(reinterpret_cast<float*>(r1), gs, gs_rcp, out[2 * h + 1], sfb[2 * h + 1], mk[2 * h + 1], nvq);
          }
        } else {
          // NVFP4: a chunk already holds two independent 16-groups and its arithmetic needs more registers (two scales,
          // e4m3 round trips); two chunks at a time spilled (45 local loads / stores per tile), so one chunk per round trip
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            uint32_t r0[32];
            tmem_ld_32x32b_x32(lane_taddr + c * 32, r0);
            tmem_ld_wait();
            if (c == 3) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty_bar(grp));
            }
            chunk_quantise<NV, METHOD, MASK>(reinterpret_cast<float*>(r0), gs, gs_rcp, out[c], sfb[c], mk[c], nvq);
          }
        }

        // ---- codes: row-per-thread -> padded staging -> 512 contiguous bytes per warp store
        const int64_t wchunk0 = (tile * kTcTileRows + q * 32) * 4;     // first chunk of this warp's 32 rows
#pragma unroll
        for (int c = 0; c < 4; ++c)
          *reinterpret_cast<uint4*>(ostage + lane * kTcOutRowBytes + c * 16) = make_uint4(out[c][0], out[c][1], out[c][2], out[c][3]);
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int u = c * 32 + lane;
          const uint4 w = *reinterpret_cast<const uint4*>(ostage + (u >> 2) * kTcOutRowBytes + (u & 3) * 16);
          if (wchunk0 + u < p.n_chunks) p.q[wchunk0 + u] = w;
        }
        __syncwarp();

        // ---- scales (+ clip mask) of this thread's row
        const int64_t chunk0 = wchunk0 + lane * 4;
        if (chunk0 < p.n_chunks) {
          if constexpr (MASK) {
            if (p.mask) *reinterpret_cast<uint4*>(p.mask + chunk0) = make_uint4(mk[0], mk[1], mk[2], mk[3]);
          }
          if constexpr (!NV) {
            const uint32_t word = sfb[0] | (sfb[1] << 8) | (sfb[2] << 16) | (sfb[3] << 24);
            if (p.sf_rm) *reinterpret_cast<uint32_t*>(p.sf_rm + chunk0) = word;
            if (p.sf_blk) {
              const uint32_t cols = (uint32_t)p.cols;
              if (row_cells) {
                const uint32_t r = (uint32_t)chunk0 / cols, c = (uint32_t)chunk0 - r * cols;
                *reinterpret_cast<uint32_t*>(p.sf_blk + sf_blocked_offset(r, c, p.padded_cols)) = word;
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint32_t n = (uint32_t)chunk0 + i, r = n / cols, c = n - r * cols;
                  p.sf_blk[sf_blocked_offset(r, c, p.padded_cols)] = (uint8_t)sfb[i];
                }
              }
            }
          } else {
            const uint32_t lo = sfb[0] | (sfb[1] << 16), hi = sfb[2] | (sfb[3] << 16);
            if (p.sf_rm) *reinterpret_cast<uint2*>(p.sf_rm + chunk0 * 2) = make_uint2(lo, hi);
            if (p.sf_blk) {
              const uint32_t cols = (uint32_t)p.cols, g0 = (uint32_t)chunk0 * 2u;
              if (row_cells) {
                const uint32_t r = g0 / cols, c = g0 - r * cols;
                uint8_t* cell = p.sf_blk + sf_blocked_offset(r, c, p.padded_cols);
                *reinterpret_cast<uint32_t*>(cell) = lo;
                *reinterpret_cast<uint32_t*>(cell + 512) = hi;
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  const uint32_t g = g0 + 2u * i, r = g / cols, c = g - r * cols;   // c is even: one 4-byte cell
                  *reinterpret_cast<uint16_t*>(p.sf_blk + sf_blocked_offset(r, c, p.padded_cols)) = (uint16_t)sfb[i];
                }
              }
            }
          }
        }
      }
    }
  }

  zero_fill_sf_padding(p, (int64_t)blockIdx.x * kTcThreads + threadIdx.x, (int64_t)gridDim.x * kTcThreads);

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc<1>(tmem_base, 512);
  }
}

bool quantize_tc_eligible(const QuantParams& p, int had, bool nv) {
  if (!(had == 32 || had == 64 || had == 128 || (nv && had == 16))) return false;
  if ((p.n_chunks & 3) != 0) return false;                         // numel % 128 == 0: whole 128-element rows
  if (p.n_chunks >= ((int64_t)1 << 31)) return false;              // 32-bit scale index math, TMA coordinate range
  if (((uintptr_t)p.rot & 15) != 0 || ((uintptr_t)p.sf_rm & 7) != 0 || ((uintptr_t)p.mask & 15) != 0) return false;
  return true;
}

template <bool NV, int METHOD, bool MASK>
static int launch_tc(const QuantParams& p, int had, cudaStream_t stream) {
  auto kern = quantize_tc_kernel<NV, METHOD, MASK>;
  static std::atomic<unsigned long long> smem_attr_done{0};   // per instantiation, one bit per device
  if (int rc_attr = ensure_dynamic_smem(kern, kTcSmem, smem_attr_done)) return rc_attr;
  const int64_t rows = p.n_chunks / 4;
  const int64_t n_tiles = ceil_div(rows, kTcTileRows);
  CUtensorMap tm;
  int rc = make_x128_tmap(&tm, p.x, rows);
  if (rc) return rc;
  CUtensorMap tr;
  rc = make_rot_tmap(&tr, p.rot, had);
  if (rc) return rc;
  int64_t ctas = n_tiles < num_sms() ? n_tiles : num_sms();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = kTcSmem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  cfg.attrs = attr;
  cfg.numAttrs = pdl_attribute(attr);
  B200Q_CUDA(cudaLaunchKernelEx(&cfg, kern, tm, tr, p, had, n_tiles));
  return 0;
}

int launch_quantize_tc(const QuantParams& p, int had, bool nv, int method, cudaStream_t stream) {
  if (nv) {
    if (method == B200Q_METHOD_QUEST) return launch_tc<true, B200Q_METHOD_QUEST, false>(p, had, stream);
    return launch_tc<true, B200Q_METHOD_ABSMAX, false>(p, had, stream);
  }
  if (method == B200Q_METHOD_QUEST) {
    if (p.mask) return launch_tc<false, B200Q_METHOD_QUEST, true>(p, had, stream);
    return launch_tc<false, B200Q_METHOD_QUEST, false>(p, had, stream);
  }
  return launch_tc<false, B200Q_METHOD_ABSMAX, false>(p, had, stream);
}

}  // namespace b200q
