// Block-scaled FP4 x FP4 -> bf16 GEMM for sm_100a, hand-written tcgen05 / TMEM / TMA.
//
// Replaces the reference's CUTLASS CollectiveBuilder kernels
// (qutlass/csrc/gemm.cu:174-326: matmul_host_mxf4_bf16_tn / matmul_host_nvf4_bf16_tn).
//
//   D[M,N] = bf16_rne( alpha * sum_k (A[m,k] * SFA[m,k/g]) * (B[n,k] * SFB[n,k/g]) ),  fp32 accumulation in TMEM
//   A [M,K/2], B [N,K/2] packed e2m1 (K-major); SFA/SFB in the cuBLAS block-scaled layout
//   (512-B blocks = 128 rows x 4 scales, K-blocks fastest), g = 32 (ue8m0) or 16 (ue4m3).
//
// Kernel shape (persistent, warp-specialised, 192 threads, 1 CTA / SM):
//   warp 0     TMA producer: per k-tile (256 K-elements = one 128-byte swizzle row) loads the A tile
//              (128 rows), this CTA's part of the B tile, and the SFA / SFB blocks into a shared-memory
//              stage; completion via mbarrier complete_tx.
//   warp 1     MMA issuer (one elected lane): copies the stage's scale blocks smem -> TMEM with
//              tcgen05.cp (32x128b.warpx4), then 4 x tcgen05.mma.kind::mxf4[nvf4].block_scale (K = 64 each);
//              tcgen05.commit releases the stage and, after the last k-tile, publishes the accumulator.
//              With kCtaGroup == 2 the CTA pair computes a 256 x BN tile (cta_group::2), each CTA holding
//              128 rows of A and half of the B rows; the peer CTA's warp 1 relays "my stage landed" to the
//              leader's barrier.
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns, alpha (device scalar) in fp32, RNE to bf16,
//              vectorised 16-byte global stores; double-buffered accumulators overlap it with the next tile.
//
// Bit-exactness (reference tests demand out == bf16(fp64 matmul)): a single fp32 accumulation chain
// per output over all of K (no split-K), alpha applied once in fp32, one RNE to bf16.
#include "common.cuh"
#include "ptx.cuh"

#include <cuda.h>
#include <mutex>

namespace b200q {
using namespace ptx;

constexpr int BM = 128;          // rows of A per CTA
constexpr int BK_BYTES = 128;    // one k-tile = 256 e2m1 = 128 bytes per row (one 128B-swizzle row)
constexpr int BK = 256;
constexpr int kGemmThreads = 192;
constexpr int kSmemBudget = 227 * 1024;

__host__ __device__ constexpr int cgcd(int a, int b) { return b == 0 ? a : cgcd(b, a % b); }

template <int kCtaGroup, int BN, bool kNV>
struct GemmCfg {
  static constexpr int SFKB = kNV ? 4 : 2;                       // 512-B scale blocks per 128 rows per k-tile
  static constexpr int G = cgcd(BN, 128);
  static constexpr int NB = (128 - G + BN + 127) / 128;          // SFB row-blocks a tile can touch
  static constexpr int B_ROWS = BN / kCtaGroup;                  // B rows this CTA stages
  static constexpr int A_BYTES = BM * BK_BYTES;
  static constexpr int B_BYTES = B_ROWS * BK_BYTES;
  static constexpr int SFA_BYTES = SFKB * 512;
  static constexpr int SFB_BYTES = NB * SFKB * 512;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES + SFA_BYTES + SFB_BYTES;
  static constexpr int SFA_COLS = SFKB * 4;
  static constexpr int SFB_COLS = SFKB * 4 * NB;
  static constexpr int ACC_STAGES = (2 * BN + SFA_COLS + SFB_COLS <= 512) ? 2 : 1;
  static constexpr int TMEM_USED = ACC_STAGES * BN + SFA_COLS + SFB_COLS;
  static constexpr int TMEM_COLS = TMEM_USED <= 32 ? 32 : TMEM_USED <= 64 ? 64 : TMEM_USED <= 128 ? 128 : TMEM_USED <= 256 ? 256 : 512;
  static constexpr int BAR_BYTES = 1024;
  static constexpr int STAGES_RAW = (kSmemBudget - BAR_BYTES - 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + BAR_BYTES + 1024;  // +1024 alignment slack
  static_assert(TMEM_USED <= 512, "TMEM overflow");
  static_assert(STAGES >= 2, "not enough shared memory for 2 stages");
  static_assert(BN % 32 == 0 && BN >= 32 && BN <= 256, "BN must be a multiple of 32 in [32, 256]");
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024-byte alignment for the 128B swizzle");
};

struct GemmParams {
  const uint8_t* sfa;
  const uint8_t* sfb;
  const float* alpha;
  __nv_bfloat16* d;
  int M, N, K;
  int tiles_m;        // ceil(M / (BM * cta_group))   (cluster tiles along M)
  int tiles_n;        // ceil(N / BN)
  int k_tiles;        // ceil(K / 256)
  int sf_col_blocks;  // ceil(K / group / 4): 512-B blocks per row-block in SFA / SFB
  int sfa_row_blocks; // ceil(M / 128)
  int sfb_row_blocks; // ceil(N / 128)
};

template <int kCtaGroup, int BN, bool kNV>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_fp4_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                const GemmParams p) {
  using Cfg = GemmCfg<kCtaGroup, BN, kNV>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int ACC = Cfg::ACC_STAGES;
  constexpr int SFKB = Cfg::SFKB;
  constexpr int NB = Cfg::NB;

  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (128B swizzle atoms); identical offset in both CTAs of a pair
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier map (8 bytes each)
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto peer_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };          // leader only: peer CTA's stage landed
  auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * STAGES + ACC + a); };
  const uint32_t tmem_slot = bar_base + 8u * (3 * STAGES + 2 * ACC);
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 8 * (3 * STAGES + 2 * ACC));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCtaGroup == 2) ? cluster_ctarank() : 0u;
  const bool is_leader = cta_rank == 0;
  const int cluster_id = blockIdx.x / kCtaGroup;
  const int num_clusters = gridDim.x / kCtaGroup;
  const int total_tiles = p.tiles_m * p.tiles_n;

  // ------------------------------------------------------------------ setup
  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&tmap_a);
    prefetch_tensormap(&tmap_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(peer_bar(s), 1);
    }
    for (int a = 0; a < ACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4 * kCtaGroup);   // 4 epilogue warps per CTA arrive on the leader's barrier
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
  const uint32_t tmem_base = *tmem_slot_gen;
  const uint32_t tmem_sfa = tmem_base + ACC * BN;
  const uint32_t tmem_sfb = tmem_sfa + Cfg::SFA_COLS;

  // ------------------------------------------------------------------ roles
  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
        const int m0 = (tm * kCtaGroup + (int)cta_rank) * BM;            // this CTA's A rows
        const int n0 = tn * BN;
        const int nb0 = n0 + (int)cta_rank * Cfg::B_ROWS;                // this CTA's B rows
        const int sfa_rb = m0 / 128;
        const int sfb_rb0 = n0 / 128;
        for (int kt = 0; kt < p.k_tiles; ++kt) {
          mbar_wait(empty_bar(stage), phase ^ 1, 1);
          const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint32_t ssfa = sb + Cfg::B_BYTES;
          const uint32_t ssfb = ssfa + Cfg::SFA_BYTES;
          // scale blocks available for this k-tile (K tail: fewer than SFKB)
          int kb_avail = p.sf_col_blocks - kt * SFKB;
          if (kb_avail > SFKB) kb_avail = SFKB;
          uint32_t tx = Cfg::A_BYTES + Cfg::B_BYTES;
          const bool sfa_ok = sfa_rb < p.sfa_row_blocks;
          if (sfa_ok) tx += kb_avail * 512;
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
            if (sfb_rb0 + nb < p.sfb_row_blocks) tx += kb_avail * 512;
          mbar_arrive_expect_tx(full_bar(stage), tx);
          tma_load_2d(sa, &tmap_a, full_bar(stage), kt * BK_BYTES, m0);
          tma_load_2d(sb, &tmap_b, full_bar(stage), kt * BK_BYTES, nb0);
          if (sfa_ok)
            bulk_load_1d(ssfa, p.sfa + ((int64_t)sfa_rb * p.sf_col_blocks + (int64_t)kt * SFKB) * 512, kb_avail * 512,
                         full_bar(stage));
#pragma unroll
          for (int nb = 0; nb < NB; ++nb)
            if (sfb_rb0 + nb < p.sfb_row_blocks)
              bulk_load_1d(ssfb + nb * SFKB * 512,
                           p.sfb + ((int64_t)(sfb_rb0 + nb) * p.sf_col_blocks + (int64_t)kt * SFKB) * 512,
                           kb_avail * 512, full_bar(stage));
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (is_leader) {
      // ===================== MMA issuer =====================
      if (lane == 0) {
        constexpr uint32_t idesc_base = make_idesc_fp4(BM * kCtaGroup, BN, !kNV);
        int stage = 0;
        uint32_t phase = 0;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
          const int tn = tile / p.tiles_m;
          const int n0 = tn * BN;
          const uint32_t sfb_shift = (uint32_t)((n0 % 128) / 32);     // column shift inside the first SFB block
          mbar_wait<kCtaGroup == 2>(tempty_bar(acc), acc_phase ^ 1, 2);
          tc_fence_after();
          const uint32_t tmem_acc = tmem_base + acc * BN;
          for (int kt = 0; kt < p.k_tiles; ++kt) {
            mbar_wait(full_bar(stage), phase, 3);
            if constexpr (kCtaGroup == 2) mbar_wait<true>(peer_bar(stage), phase, 4);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * Cfg::STAGE_BYTES;
            const uint32_t sb = sa + Cfg::A_BYTES;
            const uint32_t ssfa = sb + Cfg::B_BYTES;
            const uint32_t ssfb = ssfa + Cfg::SFA_BYTES;
            // scales: smem -> TMEM (each 512-B block -> 4 columns, replicated over the 4 lane quarters)
#pragma unroll
            for (int b = 0; b < SFKB; ++b)
              tmem_cp_32x128b_warpx4<kCtaGroup>(tmem_sfa + b * 4, make_smem_desc(ssfa + b * 512, 0, 128, kLayoutNone));
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
              for (int b = 0; b < SFKB; ++b)
                tmem_cp_32x128b_warpx4<kCtaGroup>(tmem_sfb + b * (4 * NB) + nb * 4,
                                                  make_smem_desc(ssfb + (nb * SFKB + b) * 512, 0, 128, kLayoutNone));
            // 4 MMAs of K = 64 (32 bytes along the swizzled row each)
            int kblocks = (p.K - kt * BK + 63) / 64;
            if (kblocks > 4) kblocks = 4;
            const uint64_t adesc = make_smem_desc(sa, 16, 1024, kLayoutSw128);
            const uint64_t bdesc = make_smem_desc(sb, 16, 1024, kLayoutSw128);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb) {
              if (kb < kblocks) {
                const uint32_t chunk = kNV ? kb : (kb >> 1);
                const uint32_t sf_id = kNV ? 0u : (uint32_t)((kb & 1) * 2);
                mma_fp4_block_scaled<kCtaGroup, kNV>(tmem_acc, adesc + (uint64_t)(kb * 2), bdesc + (uint64_t)(kb * 2),
                                                     idesc_with_sf_id(idesc_base, sf_id, sf_id), tmem_sfa + chunk * 4,
                                                     tmem_sfb + chunk * (4 * NB) + sfb_shift,
                                                     (kt > 0 || kb > 0) ? 1u : 0u);
              }
            }
            tc_commit<kCtaGroup>(empty_bar(stage));              // stage free once these MMAs have read it
            if (kt == p.k_tiles - 1) tc_commit<kCtaGroup>(tfull_bar(acc));   // accumulator complete
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
          if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
        }
      }
    } else {
      // ===================== peer relay (2-CTA only) =====================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
          for (int kt = 0; kt < p.k_tiles; ++kt) {
            mbar_wait(full_bar(stage), phase, 5);
            mbar_arrive_cluster(mapa(peer_bar(stage), 0));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const float alpha = __ldg(p.alpha);
    const bool vec_ok = (p.N % 8) == 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    const uint32_t tempty_leader = (kCtaGroup == 2) ? mapa(tempty_bar(0), 0) : tempty_bar(0);
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int tm = tile % p.tiles_m, tn = tile / p.tiles_m;
      const int m0 = (tm * kCtaGroup + (int)cta_rank) * BM;
      const int n0 = tn * BN;
      const int row = m0 + q * 32 + lane;
      mbar_wait(tfull_bar(acc), acc_phase, 6);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
      __nv_bfloat16* drow = p.d + (int64_t)row * p.N + n0;
#pragma unroll 1
      for (int c = 0; c < BN; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32b_x32(taddr + c, r);
        tmem_ld_wait();
        if (c + 32 >= BN) {
          // all of this warp's TMEM reads for the tile are done -> release the accumulator
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (kCtaGroup == 2) mbar_arrive_cluster(tempty_leader + 8u * acc);
            else mbar_arrive(tempty_bar(acc));
          }
        }
        if (row < p.M) {
          uint32_t packed[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float lo = __uint_as_float(r[2 * i]) * alpha;
            const float hi = __uint_as_float(r[2 * i + 1]) * alpha;
            __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
            packed[i] = *reinterpret_cast<uint32_t*>(&h2);
          }
          if (vec_ok) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              if (n0 + c + v * 8 < p.N)
                *reinterpret_cast<uint4*>(drow + c + v * 8) =
                    make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (n0 + c + i < p.N) {
                const uint32_t w = packed[i >> 1];
                reinterpret_cast<uint16_t*>(drow)[c + i] = (uint16_t)((i & 1) ? (w >> 16) : (w & 0xffffu));
              }
            }
          }
        }
      }
      if (++acc == ACC) { acc = 0; acc_phase ^= 1; }
    }
  }

  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  if constexpr (kCtaGroup == 2) cluster_sync(); else __syncthreads();
  if (warp == 1) {
    __syncwarp();   // .sync.aligned: the issuing lane must have reconverged with its warp
    tmem_dealloc<kCtaGroup>(tmem_base, Cfg::TMEM_COLS);
  }
}

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

// [rows, row_bytes] uint8 tensor, box = [box_rows, 128 bytes], 128B swizzle, zero fill out of bounds
static int make_operand_tmap(CUtensorMap* tm, const void* ptr, int64_t rows, int64_t row_bytes, int box_rows) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled entry point not available (driver too old?)");
    return B200Q_ECUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)row_bytes, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)row_bytes};
  cuuint32_t box[2] = {(cuuint32_t)BK_BYTES, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld row_bytes=%lld box_rows=%d)", (int)r,
              (long long)rows, (long long)row_bytes, box_rows);
    return B200Q_ECUDA;
  }
  return 0;
}

template <int kCtaGroup, int BN, bool kNV>
static int launch_gemm(const void* A, const void* B, const void* SFA, const void* SFB, const float* alpha, void* D,
                       int M, int N, int K, cudaStream_t stream) {
  using Cfg = GemmCfg<kCtaGroup, BN, kNV>;
  auto kern = gemm_fp4_kernel<kCtaGroup, BN, kNV>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    B200Q_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  CUtensorMap ta, tb;
  int rc = make_operand_tmap(&ta, A, M, K / 2, BM);
  if (rc) return rc;
  rc = make_operand_tmap(&tb, B, N, K / 2, Cfg::B_ROWS);
  if (rc) return rc;
  GemmParams p;
  const int group = kNV ? 16 : 32;
  p.sfa = (const uint8_t*)SFA;
  p.sfb = (const uint8_t*)SFB;
  p.alpha = alpha;
  p.d = (__nv_bfloat16*)D;
  p.M = M; p.N = N; p.K = K;
  p.tiles_m = (int)ceil_div(M, BM * kCtaGroup);
  p.tiles_n = (int)ceil_div(N, BN);
  p.k_tiles = (int)ceil_div(K, BK);
  p.sf_col_blocks = (int)ceil_div(ceil_div(K, group), 4);
  p.sfa_row_blocks = (int)ceil_div(M, 128);
  p.sfb_row_blocks = (int)ceil_div(N, 128);
  const int total = p.tiles_m * p.tiles_n;
  int clusters = num_sms() / kCtaGroup;
  if (clusters > total) clusters = total;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(clusters * kCtaGroup));
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attrs[1];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCtaGroup;
  attrs[0].val.clusterDim.y = 1;
  attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  B200Q_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, p));
  return 0;
}

template <bool kNV>
static int dispatch_cfg(int cta_group, int block_n, const void* A, const void* B, const void* SFA, const void* SFB,
                        const float* alpha, void* D, int M, int N, int K, cudaStream_t s) {
#define B200Q_CASE(CG, BNV) \
  if (cta_group == CG && block_n == BNV) return launch_gemm<CG, BNV, kNV>(A, B, SFA, SFB, alpha, D, M, N, K, s);
  B200Q_CASE(1, 64)
  B200Q_CASE(1, 128)
  B200Q_CASE(1, 256)
  B200Q_CASE(2, 128)
  B200Q_CASE(2, 256)
#undef B200Q_CASE
  set_error("unsupported GEMM configuration cta_group=%d block_n=%d", cta_group, block_n);
  return B200Q_EINVAL;
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_gemm_fp4_cfg(const void* A, const void* B, const void* SFA, const void* SFB,
                                  const float* alpha_dev, void* D_bf16, int M, int N, int K, int kind, int cta_group,
                                  int block_n, b200q_stream_t stream) {
  int rc = check_device_sm100();
  if (rc) return rc;
  B200Q_REQUIRE(A && B && SFA && SFB && alpha_dev && D_bf16, "null pointer argument");
  B200Q_REQUIRE(kind == B200Q_KIND_MXF4 || kind == B200Q_KIND_NVF4, "invalid kind %d", kind);
  B200Q_REQUIRE(M > 0 && N > 0 && K > 0, "M, N, K must be positive (got %d, %d, %d)", M, N, K);
  B200Q_REQUIRE(K % 32 == 0, "K (%d) must be a multiple of 32", K);
  B200Q_REQUIRE((((uintptr_t)A | (uintptr_t)B | (uintptr_t)SFA | (uintptr_t)SFB) & 15) == 0,
                "A, B, SFA, SFB must be 16-byte aligned");
  B200Q_REQUIRE(((uintptr_t)D_bf16 & 15) == 0 || (N % 8) != 0, "D must be 16-byte aligned");
  if (cta_group == 0 || block_n == 0) {
    // heuristic: small M streams weights with many narrow tiles; large M uses the widest tile
    if (M <= 128) { cta_group = 1; block_n = 64; }
    else { cta_group = 1; block_n = 128; }
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (kind == B200Q_KIND_NVF4) return dispatch_cfg<true>(cta_group, block_n, A, B, SFA, SFB, alpha_dev, D_bf16, M, N, K, s);
  return dispatch_cfg<false>(cta_group, block_n, A, B, SFA, SFB, alpha_dev, D_bf16, M, N, K, s);
}

extern "C" int b200q_gemm_fp4(const void* A, const void* B, const void* SFA, const void* SFB, const float* alpha_dev,
                              void* D_bf16, int M, int N, int K, int kind, b200q_stream_t stream) {
  return b200q_gemm_fp4_cfg(A, B, SFA, SFB, alpha_dev, D_bf16, M, N, K, kind, 0, 0, stream);
}
