// Per-warp-tile building blocks of the fused rotate + quantise path, shared by the standalone streaming kernel
// (quantize.cu) and by the quantiser warps of the fused quantise+GEMM kernel (gemm_fp4.cu).
// A warp-tile = 32 chunks of 32 bf16 (2 KB in, 512 B of e2m1 + scales out); lane L owns chunk L.
#pragma once
#include "common.cuh"
#include <cuda_fp8.h>

namespace b200q {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;

struct QuantParams {
  const uint4* x;          // bf16 input viewed as 16-byte units (8 bf16)
  const __nv_bfloat16* rot;
  uint4* q;                // one uint4 (32 e2m1) per chunk
  uint8_t* sf_rm;          // may be null
  uint8_t* sf_blk;         // may be null
  uint32_t* mask;          // may be null
  const float* gs;         // NV only
  int64_t n_chunks;        // numel / 32
  int64_t n_tiles;         // ceil(n_chunks / 32)
  int64_t cols;            // scales per row (row_len / group)
  int64_t padded_cols;
  int64_t rows;
  int64_t padded_rows;
  int trust_hadamard;      // caller asserted R = c * Sylvester-Hadamard (B200Q_ROT_TRUSTED_HADAMARD)
  int nv_sm100_codes;      // NVFP4 abs_max H = 128: codes from the UNROUNDED scale, like the reference's sm_100 kernel (b200q.h)
};

__device__ __forceinline__ float rcp_approx_ftz(float a) {
  float b;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(a));
  return b;
}

// 8 floats -> 8 e2m1 codes in one 32-bit word; element 2i in the low nibble of byte i.
// (same instruction the reference uses: epilogue_quant.h:78-97)
__device__ __forceinline__ uint32_t cvt8_e2m1(const float* a) {
  uint32_t val;
  asm volatile(
      "{\n"
      ".reg .b8 b0, b1, b2, b3;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b0, %2, %1;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b1, %4, %3;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b2, %6, %5;\n"
      "cvt.rn.satfinite.e2m1x2.f32 b3, %8, %7;\n"
      "mov.b32 %0, {b0, b1, b2, b3};\n"
      "}"
      : "=r"(val)
      : "f"(a[0]), "f"(a[1]), "f"(a[2]), "f"(a[3]), "f"(a[4]), "f"(a[5]), "f"(a[6]), "f"(a[7]));
  return val;
}

// Packed fp32x2 arithmetic (FADD2 / FMUL2 / FFMA2 on sm_100): the 32 values of a chunk are treated as 16 pairs
// (v[i], v[i+16]); every packed op below uses that pairing so the register allocator keeps them adjacent.
#define B200Q_PK(v, i) make_float2((v)[(i)], (v)[(i) + 16])
#define B200Q_UNPK(v, i, f) \
  do {                      \
    const float2 _t = (f);  \
    (v)[(i)] = _t.x;        \
    (v)[(i) + 16] = _t.y;   \
  } while (0)

// In-register Walsh-Hadamard butterflies over index bits [0, log2(N)) of v[0..31].
template <int N>
__device__ __forceinline__ void fwht_inreg(float* v) {
#pragma unroll
  for (int s = 1; s < 16 && s < N; s <<= 1) {           // bits 0..3: packed, both 16-halves at once
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      if ((i & s) == 0) {
        const float2 a = B200Q_PK(v, i), b = B200Q_PK(v, i | s);
        B200Q_UNPK(v, i, __fadd2_rn(a, b));
        B200Q_UNPK(v, i | s, __fadd2_rn(a, make_float2(-b.x, -b.y)));
      }
    }
  }
  if constexpr (N >= 32) {                                // bit 4: within each pair
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float a = v[i], b = v[i + 16];
      v[i] = a + b;
      v[i + 16] = a - b;
    }
  }
}

// Butterfly stage across lanes (index bit >= 5 lives in the lane id).
__device__ __forceinline__ void fwht_lane_stage(float* v, int lane_bit) {
  // lower lane: v + o, upper lane: o - v  ==  fma(sign, v, o) with sign = +-1 (exact)
  const float sg = (threadIdx.x & lane_bit) ? -1.0f : 1.0f;
  const float2 sign = make_float2(sg, sg);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 o = make_float2(__shfl_xor_sync(0xffffffffu, v[i], lane_bit),
                                 __shfl_xor_sync(0xffffffffu, v[i + 16], lane_bit));
    B200Q_UNPK(v, i, __ffma2_rn(sign, B200Q_PK(v, i), o));
  }
}

// 4 coalesced 16-byte units per lane -> XOR-swizzled per-warp staging (2 KB) -> lane L's contiguous 32-element chunk as
// fp32 in v[0..31] (bf16 -> fp32 is a 16-bit shift).
__device__ __forceinline__ void tile_stage_unpack(const uint4 (&ld)[4], uint4* stage, int lane, float* v) {
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = i * 8 + (lane >> 2), j = lane & 3;
      stage[t * 4 + (j ^ ((t >> 1) & 3))] = ld[i];
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint4 w = stage[lane * 4 + (j ^ ((lane >> 1) & 3))];
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        v[j * 8 + e * 2 + 0] = __uint_as_float(ww[e] << 16);
        v[j * 8 + e * 2 + 1] = __uint_as_float(ww[e] & 0xffff0000u);
      }
    }
}

// Hadamard rotation of every HAD-group of the warp-tile: in-register butterflies over the lane's 32 values, warp-shuffle
// butterflies for H = 64 / 128, then the scale c = R[0][0] (one fp32 rounding, like the oracle's "kernel" arithmetic).
template <int HAD>
__device__ __forceinline__ void tile_rotate_hadamard(float* v, float c_scale) {
  fwht_inreg<(HAD < 32 ? HAD : 32)>(v);
  if constexpr (HAD >= 64) fwht_lane_stage(v, 1);
  if constexpr (HAD >= 128) fwht_lane_stage(v, 2);
  const float2 c2 = make_float2(c_scale, c_scale);
#pragma unroll
  for (int i = 0; i < 16; ++i) B200Q_UNPK(v, i, __fmul2_rn(B200Q_PK(v, i), c2));
}

// Per-group scale and e2m1 conversion of ONE rotated 32-element chunk held in v[0..31] (scaled in place): 32 codes in
// out[0..3], the scale byte(s) in sf_bytes (MX: 1 byte, NV: 2 bytes little-endian), the clip mask word.
// nvq (NVFP4 abs_max only; set for Hadamard-128 unless B200Q_NV_ORACLE_CODES): the codes are computed with the scale BEFORE
// its e4m3 rounding, like the reference's sm_100-only Hadamard-128 kernel
// (sm100_visitor_store_tma_warpspecialized.hpp:141-148,567-591); the stored scale byte is the rounded one either way.
template <bool NV, int METHOD, bool MASK>
__device__ __forceinline__ void chunk_quantise(float* v, float gs, float gs_rcp, uint32_t (&out)[4], uint32_t& sf_bytes,
                                               uint32_t& mask_word, bool nvq = false) {
    mask_word = 0;
    sf_bytes = 0;
    if constexpr (!NV) {
      float scale;
      if constexpr (METHOD == B200Q_METHOD_QUEST) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          s1 += v[i];
          s2 = fmaf(v[i], v[i], s2);
        }
        const float mean = s1 / 32.f;
        const float var = fmaf(-mean, mean, s2 / 32.f);
        scale = 1.0f;
        if (var >= 0.f) scale = (float)((double)sqrtf(var) * (2.92247856 / 6.) + 1e-8);
      } else {
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(v[i]));
        scale = amax + 1e-8f;
      }
      const uint32_t e = (__float_as_uint(scale) >> 23) & 0xffu;   // floor to 2^(e-127)
      sf_bytes = e;
      // exact 1 / 2^(e-127)
      float inv = (e >= 254u) ? __uint_as_float(0x00400000u >> (e - 254u)) : __uint_as_float((254u - e) << 23);
      if constexpr (METHOD == B200Q_METHOD_ABSMAX) inv *= 3.0f;   // (x / 2^e) * 3 == x * (3 / 2^e): one rounding either way
      {
        const float2 inv2 = make_float2(inv, inv);
#pragma unroll
        for (int i = 0; i < 16; ++i) B200Q_UNPK(v, i, __fmul2_rn(B200Q_PK(v, i), inv2));
      }
    } else {
      float os[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float* vv = v + 16 * h;
        float out_scale;
        uint8_t sfb;
        if constexpr (METHOD == B200Q_METHOD_QUEST) {
          float s1 = 0.f, s2 = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            s1 += vv[i];
            s2 = fmaf(vv[i], vv[i], s2);
          }
          const float r16 = 0.0625f;
          const float mean = s1 * r16;
          const float scale = (float)((double)sqrtf(fmaf(-mean, mean, s2 * r16)) * (2.92247856 / 6.) + 1e-8);
          const __nv_fp8_e4m3 t(scale);
          sfb = *reinterpret_cast<const uint8_t*>(&t);
          const float sq = float(t);
          out_scale = (sq > 0.f) ? rcp_approx_ftz(sq) : 0.f;
        } else {
          float amax = 0.f;
#pragma unroll
          for (int i = 0; i < 16; ++i) amax = fmaxf(amax, fabsf(vv[i]));
          float sfv = gs * (amax * rcp_approx_ftz(6.0f));
          const __nv_fp8_e4m3 t(sfv);
          sfb = *reinterpret_cast<const uint8_t*>(&t);
          if (!nvq) sfv = float(t);
          out_scale = (sfv != 0.f) ? rcp_approx_ftz(sfv * gs_rcp) : 0.f;
        }
        sf_bytes |= (uint32_t)sfb << (8 * h);
        os[h] = out_scale;
      }
      {
        const float2 os2 = make_float2(os[0], os[1]);   // v[i] belongs to 16-group 0, v[i+16] to group 1
#pragma unroll
        for (int i = 0; i < 16; ++i) B200Q_UNPK(v, i, __fmul2_rn(B200Q_PK(v, i), os2));
      }
    }
#pragma unroll
    for (int w = 0; w < 4; ++w) out[w] = cvt8_e2m1(v + 8 * w);
    if constexpr (MASK) {
#pragma unroll
      for (int i = 0; i < 32; ++i) mask_word |= (fabsf(v[i]) < 6.f) ? (1u << i) : 0u;
    }
}

// all global stores (codes, row-major and blocked scales, clip mask) of one quantised chunk (= flat element index / 32)
template <bool NV, bool MASK>
__device__ __forceinline__ void chunk_store(const QuantParams& p, int64_t chunk, const uint32_t (&out)[4], uint32_t sf_bytes,
                                            uint32_t mask_word);

// chunk_quantise + chunk_store
template <bool NV, int METHOD, bool MASK>
__device__ __forceinline__ void chunk_quantise_store(const QuantParams& p, float* v, int64_t chunk, float gs, float gs_rcp) {
    uint32_t out[4];
    uint32_t mask_word, sf_bytes;
    chunk_quantise<NV, METHOD, MASK>(v, gs, gs_rcp, out, sf_bytes, mask_word, p.nv_sm100_codes != 0);
    chunk_store<NV, MASK>(p, chunk, out, sf_bytes, mask_word);
}

template <bool NV, bool MASK>
__device__ __forceinline__ void chunk_store(const QuantParams& p, int64_t chunk, const uint32_t (&out)[4], uint32_t sf_bytes,
                                            uint32_t mask_word) {
    if (chunk < p.n_chunks) {
      p.q[chunk] = make_uint4(out[0], out[1], out[2], out[3]);
      if constexpr (MASK) {
        if (p.mask) p.mask[chunk] = mask_word;
      }
      if constexpr (!NV) {
        if (p.sf_rm) p.sf_rm[chunk] = (uint8_t)sf_bytes;
        if (p.sf_blk) {
          const uint32_t r = (uint32_t)chunk / (uint32_t)p.cols, c = (uint32_t)chunk - r * (uint32_t)p.cols;
          p.sf_blk[sf_blocked_offset(r, c, p.padded_cols)] = (uint8_t)sf_bytes;
        }
      } else {
        if (p.sf_rm) reinterpret_cast<uint16_t*>(p.sf_rm)[chunk] = (uint16_t)sf_bytes;
        if (p.sf_blk) {
          const uint32_t g = (uint32_t)chunk * 2u;
          const uint32_t r = g / (uint32_t)p.cols, c = g - r * (uint32_t)p.cols;   // c is even: both bytes share a 4-byte cell
          *reinterpret_cast<uint16_t*>(p.sf_blk + sf_blocked_offset(r, c, p.padded_cols)) = (uint16_t)sf_bytes;
        }
      }
    }
}

// lane L of a warp-tile owns chunk tile * 32 + L
template <bool NV, int METHOD, bool MASK>
__device__ __forceinline__ void tile_quantise_store(const QuantParams& p, float* v, int64_t tile, int lane, float gs,
                                                    float gs_rcp) {
  chunk_quantise_store<NV, METHOD, MASK>(p, v, tile * 32 + lane, gs, gs_rcp);
}

// Zero the padding of the blocked scale buffer (rows >= rows, cols >= cols) so the buffer is written completely;
// `tid` of `nthr` cooperating threads.
__device__ __forceinline__ void zero_fill_sf_padding(const QuantParams& p, int64_t tid, int64_t nthr) {
  if (!p.sf_blk) return;
  // pad rows: one 4-byte cell (4 K-scales of one row) per iteration, 32-bit index math
  const uint32_t pad_rows = (uint32_t)(p.padded_rows - p.rows);
  const uint32_t cpr = (uint32_t)(p.padded_cols >> 2);
  for (uint32_t i = (uint32_t)tid; i < pad_rows * cpr; i += (uint32_t)nthr) {
    const uint32_t r = (uint32_t)p.rows + i / cpr, c4 = i - (i / cpr) * cpr;
    *reinterpret_cast<uint32_t*>(p.sf_blk + sf_blocked_offset(r, 4 * c4, p.padded_cols)) = 0u;
  }
  const int64_t pad_cols = p.padded_cols - p.cols;
  for (int64_t i = tid; i < p.rows * pad_cols; i += nthr) {
    const int64_t r = i / pad_cols, c = p.cols + i % pad_cols;
    p.sf_blk[sf_blocked_offset(r, c, p.padded_cols)] = 0;
  }
}

// host side (quantize.cu): argument checks + the launch-independent fields of QuantParams
int fill_params(QuantParams& p, const void* x, const void* rot, void* q, void* sf_rm, void* sf_blk, int64_t numel,
                int64_t row_len, int had, int group);

// tensor-core (tcgen05) rotation kernel, quantize_tc.cu: any runtime rotation, numel % 128 == 0
bool quantize_tc_eligible(const QuantParams& p, int had, bool nv);
int launch_quantize_tc(const QuantParams& p, int had, bool nv, int method, cudaStream_t stream);

}  // namespace b200q
