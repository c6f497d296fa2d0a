// Transposing re-quantisers of the QAT backward pass for sm_100a (HBM-bound streaming kernels).
//
// Replaces qutlass/csrc/quartet_bwd_sm120.cu:237-734 (one 32-thread CTA per 8x32 wmma tile, byte-granular gathers):
//   backward_t_bf16                    bf16 [B, N, M]             -> MXFP4 of rotate(x^T): e2m1 [B, M, N/2], ue8m0 [B, M, N/32]
//   backward_qt_bf16                   MXFP4 [B, N, M/2] + scales -> MXFP4 of rotate(dq(x)^T) (scale / alpha)
//   backward_bf16_square_double_mxfp8  bf16 [m, n]                -> e4m3 [m_pad, n] with ONE ue8m0 scale per 32x32 tile,
//                                                                    written both as row scales and as column scales
//   mxfp4_transpose_mxfp8              MXFP4 [m, n/2] + scales    -> e4m3 [n, m_pad] of dq(x)^T, ue8m0 per 32 along m
//
// Data flow of the three transposing kernels (one 256-thread CTA per 128 x 128 input tile, n-tiles fastest so that the
// CTAs resident together fill whole output lines): coalesced 16-byte global loads -> shared tile kept in INPUT
// orientation -> thread (g, p) walks DOWN the columns (2p, 2p+1) of row group g -- conflict-free 4-byte (bf16x2) or
// 1-byte (e2m1x2) shared loads -- and so holds two complete 32-groups of the TRANSPOSED matrix in registers ->
// in-register Walsh-Hadamard butterflies (or a generic fp32 x @ R for any other rotation) -> abs-max scale -> hardware
// cvt (e2m1x2 / e4m3x2) -> shared output staging -> coalesced 16-byte stores, one 32-bit scale word per output row.
//
// Algorithmic bytes / element: backward_t 2 + 0.5 + 1/32 = 2.53; backward_qt 0.53 + 0.53 = 1.06;
// square_double 2 + 1 + 2/32 = 3.06; mxfp4_transpose_mxfp8 0.53 + 1 + 1/32 = 1.56.
#include "quantize_tile.cuh"
#include "backward_quant.cuh"
#include <cuda_fp16.h>

namespace b200q {

constexpr int kBwdThreads = 256;
constexpr int kBwdTile = 128;

struct BwdParams {
  const void* x;            // T: bf16 [B, N, M];  QT / TR8: packed e2m1 [B, N, M/2]
  const uint8_t* x_sf;      // QT / TR8: ue8m0 [B, N, M/32]
  const __nv_bfloat16* rot; // [32, 32] row-major [k, n]
  const float* alpha;       // QT: device scalar
  uint8_t* q;               // T/QT: e2m1 [B, M, N/2];  TR8: e4m3 [M, n_pad]
  uint8_t* sf;              // T/QT: ue8m0 [B, M, N/32];  TR8: ue8m0 [M, n_pad/32]
  int N;                    // rows of the input (the dimension that is grouped by 32 after the transpose)
  int M;                    // columns of the input (rows of the output)
  int n_valid;              // TR8: input rows that exist (rows in [n_valid, N) read as zero); otherwise == N
};

// ------------------------------------------------------------------ rotation of one 32-group held in registers
template <bool TRUST>
__device__ __forceinline__ void rotate32(float* v, float c_scale, const float* s_rot) {
  if constexpr (TRUST) {
    fwht_inreg<32>(v);
    const float2 c2 = make_float2(c_scale, c_scale);
#pragma unroll
    for (int i = 0; i < 16; ++i) B200Q_UNPK(v, i, __fmul2_rn(B200Q_PK(v, i), c2));
  } else {
    // generic xh[j] = sum_k x[k] R[k][j]; R as fp32 in shared memory (broadcast reads)
    float o[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) o[j] = 0.f;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float xk = v[k];
      const float4* r4 = reinterpret_cast<const float4*>(s_rot + k * 32);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4 r = r4[j];
        o[4 * j + 0] = fmaf(xk, r.x, o[4 * j + 0]);
        o[4 * j + 1] = fmaf(xk, r.y, o[4 * j + 1]);
        o[4 * j + 2] = fmaf(xk, r.z, o[4 * j + 2]);
        o[4 * j + 3] = fmaf(xk, r.w, o[4 * j + 3]);
      }
    }
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = o[j];
  }
}

// ------------------------------------------------------------------ tile loads (input orientation)
// bf16 tile: 128 rows x 256 B
__device__ __forceinline__ void load_tile_bf16(const BwdParams& p, const __nv_bfloat16* xb, int n0, int m0, uint8_t* s_tile) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + kBwdThreads * i;
    const int row = c >> 4, ch = c & 15;
    const int n = n0 + row, m = m0 + ch * 8;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < p.n_valid && m < p.M) v = __ldg(reinterpret_cast<const uint4*>(xb + (size_t)n * p.M + m));
    *reinterpret_cast<uint4*>(s_tile + c * 16) = v;
  }
}

// packed e2m1 tile: 128 rows x 64 B of codes + 128 x 4 scale bytes (rows / groups outside the matrix: code 0, scale 127)
__device__ __forceinline__ void load_tile_fp4(const BwdParams& p, const uint8_t* xq, const uint8_t* xs, int n0, int m0,
                                              uint8_t* s_tile, uint8_t* s_sc) {
  const int row_bytes = p.M >> 1, row_sf = p.M >> 5;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + kBwdThreads * i;
    const int row = c >> 2, ch = c & 3;
    const int n = n0 + row, m = m0 + ch * 32;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (n < p.n_valid && m < p.M) v = __ldg(reinterpret_cast<const uint4*>(xq + (size_t)n * row_bytes + (m >> 1)));
    *reinterpret_cast<uint4*>(s_tile + c * 16) = v;
  }
  if (threadIdx.x < kBwdTile) {
    const int n = n0 + threadIdx.x;
    uint32_t w = 0x7f7f7f7fu;
    if (n < p.n_valid) {
      const uint8_t* src = xs + (size_t)n * row_sf + (m0 >> 5);
      if ((row_sf & 3) == 0 && m0 + 128 <= p.M) {
        w = __ldg(reinterpret_cast<const uint32_t*>(src));
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (m0 + 32 * j < p.M) w = (w & ~(0xffu << (8 * j))) | ((uint32_t)__ldg(src + j) << (8 * j));
        }
      }
    }
    reinterpret_cast<uint32_t*>(s_sc)[threadIdx.x] = w;
  }
}

// ------------------------------------------------------------------ backward_t_bf16 / backward_qt_bf16
// compute phase: shared input tile (input orientation) -> rotated + quantised output tile in shared staging
template <bool QT, bool TRUST>
__device__ __forceinline__ void bwd_tq_compute(const uint8_t* s_tile, const uint8_t* s_sc, uint8_t* s_out, uint8_t* s_osf,
                                               const float* s_rot, float c_scale, float alpha) {
  const int g = threadIdx.x >> 6, mp = threadIdx.x & 63;   // row group (32 input rows) / column pair
  float v0[32], v1[32];
  if constexpr (QT) {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int r = 32 * g + k;
      const float2 f = e2m1x2_to_float2(s_tile[r * 64 + mp]);
      // (uint16) byte << 7 as bf16 == byte << 23 as fp32: 2^(e-127), e = 0 -> 0.0 (quartet_bwd_sm120.cu:360)
      const float sc = __uint_as_float((uint32_t)s_sc[r * 4 + (mp >> 4)] << 23);
      v0[k] = f.x * sc;
      v1[k] = f.y * sc;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(s_tile + (32 * g + k) * 256 + mp * 4);
      v0[k] = __uint_as_float(w << 16);
      v1[k] = __uint_as_float(w & 0xffff0000u);
    }
  }
  uint32_t o0[4], o1[4];
  const float c3 = QT ? __fdiv_rn(3.0f, alpha) : 0.f;       // once per tile instead of two divisions per group (backward_quant.cuh)
  rotate32<TRUST>(v0, c_scale, s_rot);
  const uint32_t e0 = quantise32_absmax<QT>(v0, alpha, o0, c3);
  rotate32<TRUST>(v1, c_scale, s_rot);
  const uint32_t e1 = quantise32_absmax<QT>(v1, alpha, o1, c3);

  // stage: output row (2 mp + col) holds 4 groups x 16 B; the 16-byte slot is XOR-swizzled with (row >> 1) & 3
  const int slot = (g ^ (mp & 3)) * 16;
  *reinterpret_cast<uint4*>(s_out + (2 * mp) * 64 + slot) = make_uint4(o0[0], o0[1], o0[2], o0[3]);
  *reinterpret_cast<uint4*>(s_out + (2 * mp + 1) * 64 + slot) = make_uint4(o1[0], o1[1], o1[2], o1[3]);
  s_osf[(2 * mp) * 4 + g] = (uint8_t)e0;
  s_osf[(2 * mp + 1) * 4 + g] = (uint8_t)e1;
}

// store phase: staged output tile -> coalesced 16-byte stores, one 32-bit scale word per output row
__device__ __forceinline__ void bwd_tq_store(const BwdParams& p, const uint8_t* s_out, const uint8_t* s_osf, int n0, int m0, int b) {
  const int out_row_bytes = p.N >> 1, out_row_sf = p.N >> 5;
  uint8_t* qb = p.q + (size_t)b * p.M * out_row_bytes;
  uint8_t* sb = p.sf + (size_t)b * p.M * out_row_sf;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int c = threadIdx.x + kBwdThreads * i;
    const int row = c >> 2, part = c & 3;
    const int m = m0 + row, n = n0 + part * 32;
    if (m < p.M && n < p.N) {
      const uint4 v = *reinterpret_cast<const uint4*>(s_out + row * 64 + ((part ^ ((row >> 1) & 3)) * 16));
      *reinterpret_cast<uint4*>(qb + (size_t)m * out_row_bytes + (n >> 1)) = v;
    }
  }
  if (threadIdx.x < kBwdTile) {
    const int m = m0 + threadIdx.x;
    if (m < p.M) {
      const uint32_t w = reinterpret_cast<const uint32_t*>(s_osf)[threadIdx.x];
      uint8_t* dst = sb + (size_t)m * out_row_sf + (n0 >> 5);
      if ((out_row_sf & 3) == 0) {
        *reinterpret_cast<uint32_t*>(dst) = w;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (n0 + 32 * j < p.N) dst[j] = (uint8_t)(w >> (8 * j));
        }
      }
    }
  }
}

template <bool QT, bool TRUST>
__global__ void __launch_bounds__(kBwdThreads) bwd_transpose_quantize_fp4_kernel(const BwdParams p) {
  __shared__ __align__(16) uint8_t s_tile[QT ? kBwdTile * 64 : kBwdTile * 256];
  __shared__ __align__(16) uint8_t s_sc[QT ? kBwdTile * 4 : 16];
  __shared__ __align__(16) uint8_t s_out[kBwdTile * 64];
  __shared__ __align__(16) uint8_t s_osf[kBwdTile * 4];
  __shared__ __align__(16) float s_rot[TRUST ? 4 : 1024];

  const int n0 = blockIdx.x * kBwdTile, m0 = blockIdx.y * kBwdTile, b = blockIdx.z;
  if constexpr (QT) {
    load_tile_fp4(p, (const uint8_t*)p.x + (size_t)b * p.N * (p.M >> 1), p.x_sf + (size_t)b * p.N * (p.M >> 5), n0, m0,
                  s_tile, s_sc);
  } else {
    load_tile_bf16(p, (const __nv_bfloat16*)p.x + (size_t)b * p.N * p.M, n0, m0, s_tile);
  }
  if constexpr (!TRUST) {
    for (int i = threadIdx.x; i < 1024; i += kBwdThreads) s_rot[i] = __bfloat162float(p.rot[i]);
  }
  const float c_scale = __bfloat162float(p.rot[0]);
  float alpha = 1.f;
  if constexpr (QT) alpha = __ldg(p.alpha);
  __syncthreads();
  bwd_tq_compute<QT, TRUST>(s_tile, s_sc, s_out, s_osf, s_rot, c_scale, alpha);
  __syncthreads();
  bwd_tq_store(p, s_out, s_osf, n0, m0, b);
}

// ---- persistent, double-buffered form (round 2, session 3): a CTA walks tiles blockIdx.x, + gridDim.x, ... (n-tiles fastest, so the
// CTAs of one sweep still fill whole output lines together) and requests tile i + 1 with cp.async (zero-filled outside the
// matrix) before it computes tile i: the one-shot form exposes the DRAM latency of every tile to its two resident CTAs
// (load -> barrier -> ~900 instructions per thread -> barrier -> store).  Same compute / store code, same bytes.
// request one input tile (bf16: 128 rows x 256 B; FP4: 128 rows x 64 B of codes + 128 x 4 scale bytes) into shared memory
template <bool FP4>
__device__ __forceinline__ void bwd_request_tile(const BwdParams& p, int b, int n0, int m0, uint8_t* s_tile, uint8_t* s_sc) {
  if constexpr (!FP4) {
    const __nv_bfloat16* xb = (const __nv_bfloat16*)p.x + (size_t)b * p.N * p.M;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = threadIdx.x + kBwdThreads * i;
      const int row = c >> 4, ch = c & 15;
      const int n = n0 + row, m = m0 + ch * 8;
      const bool ok = n < p.n_valid && m < p.M;
      cp_async16_zfill(s_tile + c * 16, ok ? (const void*)(xb + (size_t)n * p.M + m) : p.x, ok);
    }
  } else {
    const int row_bytes = p.M >> 1, row_sf = p.M >> 5;
    const uint8_t* xq = (const uint8_t*)p.x + (size_t)b * p.N * row_bytes;
    const uint8_t* xs = p.x_sf + (size_t)b * p.N * row_sf;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int c = threadIdx.x + kBwdThreads * i;
      const int row = c >> 2, ch = c & 3;
      const int n = n0 + row, m = m0 + ch * 32;
      const bool ok = n < p.n_valid && m < p.M;
      cp_async16_zfill(s_tile + c * 16, ok ? (const void*)(xq + (size_t)n * row_bytes + (m >> 1)) : p.x, ok);
    }
    if (threadIdx.x < kBwdTile) {
      // outside the matrix the codes are zero, so any scale gives 0 (zero-filled scale bytes: 0 * 0 resp. 0 * 2^-127)
      const int n = n0 + threadIdx.x;
      const uint8_t* src = xs + (size_t)n * row_sf + (m0 >> 5);
      if ((row_sf & 3) == 0 && (m0 + 128 <= p.M || n >= p.n_valid)) {
        cp_async4_zfill(s_sc + threadIdx.x * 4, n < p.n_valid ? (const void*)src : (const void*)p.x_sf, n < p.n_valid);
      } else {
        uint32_t w = 0;
        if (n < p.n_valid) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            if (m0 + 32 * j < p.M) w |= (uint32_t)__ldg(src + j) << (8 * j);
          }
        }
        reinterpret_cast<uint32_t*>(s_sc)[threadIdx.x] = w;     // plain store: visible after the barrier that follows the wait
      }
    }
  }
}

constexpr int kBwdPipeOut = kBwdTile * 64 + kBwdTile * 4;                     // staged output tile + its scale words
constexpr int bwd_pipe_smem(bool fp4, bool trust) {
  return 2 * (fp4 ? kBwdTile * 64 + kBwdTile * 4 : kBwdTile * 256) + kBwdPipeOut + (trust ? 0 : 4096);
}

template <bool QT, bool TRUST>
__global__ void __launch_bounds__(kBwdThreads) bwd_transpose_quantize_fp4_pipe_kernel(const BwdParams p, int tiles_n, int tiles_m,
                                                                                     int n_tiles) {
  extern __shared__ __align__(16) uint8_t bwd_smem[];
  constexpr int kIn = QT ? kBwdTile * 64 + kBwdTile * 4 : kBwdTile * 256;
  uint8_t* s_out = bwd_smem + 2 * kIn;
  uint8_t* s_osf = s_out + kBwdTile * 64;
  float* s_rot = reinterpret_cast<float*>(s_osf + kBwdTile * 4);
  auto coords = [&](int t, int& b, int& n0, int& m0) {
    const int tn = t % tiles_n, r = t / tiles_n;
    n0 = tn * kBwdTile;
    m0 = (r % tiles_m) * kBwdTile;
    b = r / tiles_m;
  };
  int t = blockIdx.x, b, n0, m0;
  if (t < n_tiles) {
    coords(t, b, n0, m0);
    bwd_request_tile<QT>(p, b, n0, m0, bwd_smem, bwd_smem + kBwdTile * 64);
  }
  cp_async_commit();
  if constexpr (!TRUST) {
    for (int i = threadIdx.x; i < 1024; i += kBwdThreads) s_rot[i] = __bfloat162float(p.rot[i]);
  }
  const float c_scale = __bfloat162float(p.rot[0]);
  float alpha = 1.f;
  if constexpr (QT) alpha = __ldg(p.alpha);
  int buf = 0;
  for (; t < n_tiles; t += gridDim.x, buf ^= 1) {
    coords(t, b, n0, m0);
    const int tn_next = t + gridDim.x;
    if (tn_next < n_tiles) {
      int b2, n2, m2;
      coords(tn_next, b2, n2, m2);
      uint8_t* nb = bwd_smem + (buf ^ 1) * kIn;
      bwd_request_tile<QT>(p, b2, n2, m2, nb, nb + kBwdTile * 64);
    }
    cp_async_commit();
    cp_async_wait<1>();                 // this tile's group (the one before the newest) has landed
    __syncthreads();                    // ... for every thread; also: the previous tile's stores have read s_out
    const uint8_t* tile = bwd_smem + buf * kIn;
    bwd_tq_compute<QT, TRUST>(tile, tile + kBwdTile * 64, s_out, s_osf, s_rot, c_scale, alpha);
    __syncthreads();                    // staging complete; every thread is done with buffer `buf` (refilled next iteration)
    bwd_tq_store(p, s_out, s_osf, n0, m0, b);
  }
}

// ------------------------------------------------------------------ mxfp4_transpose_mxfp8
// input rows = m of the reference (grouped by 32 after the transpose), input columns = n (rows of the output).
__device__ __forceinline__ void bwd_tr8_compute(const uint8_t* s_tile, const uint8_t* s_sc, uint8_t* s_out, uint8_t* s_osf) {
  const int g = threadIdx.x >> 6, mp = threadIdx.x & 63;
  float v0[32], v1[32];
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int r = 32 * g + k;
    const float2 f = e2m1x2_to_float2(s_tile[r * 64 + mp]);
    const uint32_t e = s_sc[r * 4 + (mp >> 4)];
    // __nv_cvt_e8m0_to_bf16raw: 2^(e-127), e = 0 -> 2^-127 (quartet_bwd_sm120.cu:652-654)
    const float sc = __uint_as_float(e == 0u ? 0x00400000u : (e << 23));
    v0[k] = f.x * sc;
    v1[k] = f.y * sc;
    a0 = fmaxf(a0, fabsf(v0[k]));
    a1 = fmaxf(a1, fabsf(v1[k]));
  }
  const uint32_t e0 = e8m0_shift7(a0), e1 = e8m0_shift7(a1);
  const float i0 = inv_pow2_of_e8m0(e0), i1 = inv_pow2_of_e8m0(e1);
  uint32_t w0[8], w1[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    w0[j] = cvt2_e4m3(v0[4 * j] * i0, v0[4 * j + 1] * i0) | (cvt2_e4m3(v0[4 * j + 2] * i0, v0[4 * j + 3] * i0) << 16);
    w1[j] = cvt2_e4m3(v1[4 * j] * i1, v1[4 * j + 1] * i1) | (cvt2_e4m3(v1[4 * j + 2] * i1, v1[4 * j + 3] * i1) << 16);
  }
  // stage: output row (2 mp + col) = 128 B (4 groups x 32 B); 16-byte slots XOR-swizzled with (row >> 1) & 7
  const int sw = mp & 7;
  uint8_t* r0 = s_out + (2 * mp) * 128;
  uint8_t* r1 = r0 + 128;
  *reinterpret_cast<uint4*>(r0 + (((2 * g) ^ sw) * 16)) = make_uint4(w0[0], w0[1], w0[2], w0[3]);
  *reinterpret_cast<uint4*>(r0 + (((2 * g + 1) ^ sw) * 16)) = make_uint4(w0[4], w0[5], w0[6], w0[7]);
  *reinterpret_cast<uint4*>(r1 + (((2 * g) ^ sw) * 16)) = make_uint4(w1[0], w1[1], w1[2], w1[3]);
  *reinterpret_cast<uint4*>(r1 + (((2 * g + 1) ^ sw) * 16)) = make_uint4(w1[4], w1[5], w1[6], w1[7]);
  s_osf[(2 * mp) * 4 + g] = (uint8_t)e0;
  s_osf[(2 * mp + 1) * 4 + g] = (uint8_t)e1;
}

__device__ __forceinline__ void bwd_tr8_store(const BwdParams& p, const uint8_t* s_out, const uint8_t* s_osf, int n0, int m0) {
  const int out_row_bytes = p.N, out_row_sf = p.N >> 5;   // N is the padded row count of the input (multiple of 128)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c = threadIdx.x + kBwdThreads * i;
    const int row = c >> 3, part = c & 7;
    const int m = m0 + row;
    if (m < p.M) {
      const uint4 v = *reinterpret_cast<const uint4*>(s_out + row * 128 + ((part ^ ((row >> 1) & 7)) * 16));
      *reinterpret_cast<uint4*>(p.q + (size_t)m * out_row_bytes + n0 + part * 16) = v;
    }
  }
  if (threadIdx.x < kBwdTile) {
    const int m = m0 + threadIdx.x;
    if (m < p.M)
      *reinterpret_cast<uint32_t*>(p.sf + (size_t)m * out_row_sf + (n0 >> 5)) = reinterpret_cast<const uint32_t*>(s_osf)[threadIdx.x];
  }
}

__global__ void __launch_bounds__(kBwdThreads) bwd_mxfp4_transpose_mxfp8_kernel(const BwdParams p) {
  __shared__ __align__(16) uint8_t s_tile[kBwdTile * 64];
  __shared__ __align__(16) uint8_t s_sc[kBwdTile * 4];
  __shared__ __align__(16) uint8_t s_out[kBwdTile * 128];
  __shared__ __align__(16) uint8_t s_osf[kBwdTile * 4];

  const int n0 = blockIdx.x * kBwdTile, m0 = blockIdx.y * kBwdTile;   // n0: input row, m0: input column
  load_tile_fp4(p, (const uint8_t*)p.x, p.x_sf, n0, m0, s_tile, s_sc);
  __syncthreads();
  bwd_tr8_compute(s_tile, s_sc, s_out, s_osf);
  __syncthreads();
  bwd_tr8_store(p, s_out, s_osf, n0, m0);
}

// persistent, double-buffered form (see bwd_transpose_quantize_fp4_pipe_kernel)
constexpr int kBwdTr8PipeSmem = 2 * (kBwdTile * 64 + kBwdTile * 4) + kBwdTile * 128 + kBwdTile * 4;
__global__ void __launch_bounds__(kBwdThreads) bwd_mxfp4_transpose_mxfp8_pipe_kernel(const BwdParams p, int tiles_n, int n_tiles) {
  extern __shared__ __align__(16) uint8_t bwd_smem[];
  constexpr int kIn = kBwdTile * 64 + kBwdTile * 4;
  uint8_t* s_out = bwd_smem + 2 * kIn;
  uint8_t* s_osf = s_out + kBwdTile * 128;
  int t = blockIdx.x;
  if (t < n_tiles) bwd_request_tile<true>(p, 0, (t % tiles_n) * kBwdTile, (t / tiles_n) * kBwdTile, bwd_smem, bwd_smem + kBwdTile * 64);
  cp_async_commit();
  int buf = 0;
  for (; t < n_tiles; t += gridDim.x, buf ^= 1) {
    const int n0 = (t % tiles_n) * kBwdTile, m0 = (t / tiles_n) * kBwdTile;
    const int t2 = t + gridDim.x;
    if (t2 < n_tiles) {
      uint8_t* nb = bwd_smem + (buf ^ 1) * kIn;
      bwd_request_tile<true>(p, 0, (t2 % tiles_n) * kBwdTile, (t2 / tiles_n) * kBwdTile, nb, nb + kBwdTile * 64);
    }
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    const uint8_t* tile = bwd_smem + buf * kIn;
    bwd_tr8_compute(tile, tile + kBwdTile * 64, s_out, s_osf);
    __syncthreads();
    bwd_tr8_store(p, s_out, s_osf, n0, m0);
  }
}

// ------------------------------------------------------------------ backward_bf16_square_double_mxfp8
// One CTA (4 warps) per 128 x 128 block; warp w owns rows 32w..32w+31 (four 32 x 32 tiles side by side).
struct SquareParams {
  const __nv_bfloat16* x;   // [m_valid, n]
  uint8_t* y;               // e4m3 [m_pad, n]
  uint8_t* row_sf;          // ue8m0 [m_pad, n/32]
  uint8_t* col_sf;          // ue8m0 [n, m_pad/32]
  int m_valid, m_pad, n;
};

__global__ void __launch_bounds__(128) bwd_square_double_mxfp8_kernel(const SquareParams p) {
  __shared__ __align__(4) uint8_t s_e[4][4];   // [row tile (warp)][column tile]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = blockIdx.y * 128, j0 = blockIdx.x * 128;
  const int half = lane >> 4, ch = lane & 15;   // a load instruction covers 2 rows x 256 B
  const int col = j0 + ch * 8;
  uint4 ld[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int row = i0 + 32 * warp + 2 * i + half;
    ld[i] = (row < p.m_valid && col < p.n) ? __ldg(reinterpret_cast<const uint4*>(p.x + (size_t)row * p.n + col))
                                           : make_uint4(0, 0, 0, 0);
  }
  // abs-max of this lane's 128 values: bf16 magnitudes order like their 15-bit patterns
  uint32_t mx = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t ww[4] = {ld[i].x, ld[i].y, ld[i].z, ld[i].w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      mx = max(mx, ww[e] & 0x7fffu);
      mx = max(mx, (ww[e] >> 16) & 0x7fffu);
    }
  }
  // lanes of one 32x32 tile: same (ch >> 2) -> reduce over lane bits 0, 1 and 4
  mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
  mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
  mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
  const uint32_t e = e8m0_shift7(__uint_as_float(mx << 16));
  const float inv = inv_pow2_of_e8m0(e);
  if ((lane & 19) == 0) s_e[warp][ch >> 2] = (uint8_t)e;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const int row = i0 + 32 * warp + 2 * i + half;
    const uint32_t ww[4] = {ld[i].x, ld[i].y, ld[i].z, ld[i].w};
    uint32_t o[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float a0 = __uint_as_float(ww[2 * h] << 16) * inv, a1 = __uint_as_float(ww[2 * h] & 0xffff0000u) * inv;
      const float a2 = __uint_as_float(ww[2 * h + 1] << 16) * inv, a3 = __uint_as_float(ww[2 * h + 1] & 0xffff0000u) * inv;
      o[h] = cvt2_e4m3(a0, a1) | (cvt2_e4m3(a2, a3) << 16);
    }
    if (row < p.m_pad && col < p.n) *reinterpret_cast<uint2*>(p.y + (size_t)row * p.n + col) = make_uint2(o[0], o[1]);
  }
  __syncthreads();
  // row scales: row i0 + t gets the 4 tile bytes of its warp-row; column scales: column j0 + t gets the 4 row-tile bytes
  const int t = threadIdx.x;
  const int n_sf = p.n >> 5, m_sf = p.m_pad >> 5;
  {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(s_e[t >> 5]);
    uint8_t* dst = p.row_sf + (size_t)(i0 + t) * n_sf + (j0 >> 5);
    if ((n_sf & 3) == 0) {
      *reinterpret_cast<uint32_t*>(dst) = w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (j0 + 32 * j < p.n) dst[j] = (uint8_t)(w >> (8 * j));
      }
    }
  }
  if (j0 + t < p.n) {
    const int tj = t >> 5;
    const uint32_t w = (uint32_t)s_e[0][tj] | ((uint32_t)s_e[1][tj] << 8) | ((uint32_t)s_e[2][tj] << 16) | ((uint32_t)s_e[3][tj] << 24);
    *reinterpret_cast<uint32_t*>(p.col_sf + (size_t)(j0 + t) * m_sf + (i0 >> 5)) = w;   // m_pad % 128 == 0
  }
}

static bool aligned16(const void* ptr) { return (reinterpret_cast<uintptr_t>(ptr) & 15) == 0; }

// backward_tc.cu: backward_t_bf16 with the transpose + rotation on the tensor cores (any runtime rotation at the streaming rate)
bool backward_t_tc_eligible(const void* x, const void* rot, const void* q, const void* sf, int size_m, int size_n, int size_b);
int launch_backward_t_tc(const void* x, const void* rot, void* q, void* sf, int size_m, int size_n, int size_b, cudaStream_t stream);
bool backward_qt_tc_eligible(const void* xq, const void* xs, const void* rot, const void* q, const void* sf, int size_m, int size_n,
                             int size_b);
int launch_backward_qt_tc(const void* xq, const void* xs, const void* rot, const float* alpha, void* q, void* sf, int size_m,
                          int size_n, int size_b, cudaStream_t stream);

// B200Q_BWD_PIPE=0/1 forces the one-shot / the persistent double-buffered form of the three transposing kernels (same bytes)
// Measured (profiles/r02_s3_bwd_bench_pipe{0,1}.jsonl, graph replay over rotating sets): backward_qt 45.9 -> 43.2 us and
// mxfp4_transpose_mxfp8 36.3 -> 33.5 us at 16384 x 4096 (41.7 -> 38.2 / 34.1 -> 29.7 at 4096 x 14336); backward_qt slower at
// 4096 x 4096 (3.5 tiles per CTA: 12.8 -> 13.5), backward_t slower everywhere (its 32 KB tiles halve the resident CTAs).  The
// kernels are bound by their ~14 instructions per element, not by latency, so the gain is modest.  Library rule: the FP4-input
// kernels from 2048 tiles on.
static bool bwd_pipe_enabled(bool fp4_input, int64_t n_tiles) {
  const int sw = env().bwd_pipe;
  if (n_tiles >= ((int64_t)1 << 31)) return false;             // 32-bit tile index in the persistent kernels
  return sw == 1 || (sw < 0 && fp4_input && n_tiles >= 2048);
}

// persistent grid: resident CTAs per SM (queried once per kernel and device) x SMs, at most one CTA per tile
template <typename Kern>
static int pipe_grid(Kern kern, int smem, std::atomic<int>* occ_cache, int n_tiles) {
  std::atomic<int>& slot = occ_cache[current_device() & 63];
  int occ = slot.load(std::memory_order_acquire);
  if (occ == 0) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, kBwdThreads, smem);
    if (e != cudaSuccess || n <= 0) {
      set_error("cudaOccupancyMaxActiveBlocksPerMultiprocessor failed: %s", cudaGetErrorString(e));
      return -1;
    }
    occ = n;
    slot.store(occ, std::memory_order_release);
  }
  const int64_t ctas = (int64_t)occ * num_sms();
  return (int)(ctas < n_tiles ? ctas : n_tiles);
}

template <bool QT, bool TRUST>
static int launch_tq_pipe(const BwdParams& p, dim3 grid, cudaStream_t stream) {
  auto kern = bwd_transpose_quantize_fp4_pipe_kernel<QT, TRUST>;
  constexpr int smem = bwd_pipe_smem(QT, TRUST);
  static std::atomic<unsigned long long> attr_done{0};
  static std::atomic<int> occ[64];
  if (int rc = ensure_dynamic_smem(kern, smem, attr_done)) return rc;
  const int64_t n_tiles64 = (int64_t)grid.x * grid.y * grid.z;
  B200Q_REQUIRE(n_tiles64 < ((int64_t)1 << 31), "problem too large for one launch");
  const int n_tiles = (int)n_tiles64;
  const int ctas = pipe_grid(kern, smem, occ, n_tiles);
  if (ctas <= 0) return B200Q_ECUDA;
  kern<<<ctas, kBwdThreads, smem, stream>>>(p, (int)grid.x, (int)grid.y, n_tiles);
  B200Q_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace b200q

using namespace b200q;

extern "C" int b200q_backward_t_bf16(const void* x_bf16, const void* rot_bf16, void* xh_e2m1, void* xh_e8m0, int size_m,
                                     int size_n, int size_b, int flags, b200q_stream_t stream) {
  int rc = check_device_sm100();
  if (rc) return rc;
  B200Q_REQUIRE(x_bf16 && rot_bf16 && xh_e2m1 && xh_e8m0, "null pointer argument");
  B200Q_REQUIRE(size_m > 0 && size_n > 0 && size_b > 0, "sizes must be positive");
  B200Q_REQUIRE(size_n % 32 == 0, "size_n (%d) must be a multiple of 32", size_n);
  B200Q_REQUIRE(size_m % 8 == 0, "size_m (%d) must be a multiple of 8", size_m);
  B200Q_REQUIRE(aligned16(x_bf16) && aligned16(xh_e2m1) && (reinterpret_cast<uintptr_t>(xh_e8m0) & 3) == 0,
                "pointers must be 16-byte aligned (scales: 4-byte)");
  BwdParams p{};
  p.x = x_bf16; p.rot = (const __nv_bfloat16*)rot_bf16; p.q = (uint8_t*)xh_e2m1; p.sf = (uint8_t*)xh_e8m0;
  p.N = size_n; p.M = size_m; p.n_valid = size_n;
  dim3 grid((unsigned)ceil_div(size_n, kBwdTile), (unsigned)ceil_div(size_m, kBwdTile), (unsigned)size_b);
  B200Q_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, "problem too large for one launch");
  {
    // B200Q_BWD_T_TC=0/1 forces the CUDA-core / the tcgen05 kernel; library rule: from 8 M elements, and a non-Hadamard rotation from
    // 1 M elements on (its CUDA-core path costs 1024 FMAs per group)
    const int sw = env().bwd_t_tc;
    const int64_t numel = (int64_t)size_m * size_n * size_b;
    const bool trusted = (flags & B200Q_ROT_TRUSTED_HADAMARD) != 0;
    if ((sw == 1 || (sw < 0 && numel >= (trusted ? ((int64_t)1 << 23) : ((int64_t)1 << 20)))) &&
        backward_t_tc_eligible(x_bf16, rot_bf16, xh_e2m1, xh_e8m0, size_m, size_n, size_b))
      return launch_backward_t_tc(x_bf16, rot_bf16, xh_e2m1, xh_e8m0, size_m, size_n, size_b, (cudaStream_t)stream);
  }
  // (a generic rotation is bound by its 1024 FMAs per group, not by latency: one-shot form only)
  if (bwd_pipe_enabled(false, (int64_t)grid.x * grid.y * grid.z) && (flags & B200Q_ROT_TRUSTED_HADAMARD)) return launch_tq_pipe<false, true>(p, grid, (cudaStream_t)stream);
  if (flags & B200Q_ROT_TRUSTED_HADAMARD)
    bwd_transpose_quantize_fp4_kernel<false, true><<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(p);
  else
    bwd_transpose_quantize_fp4_kernel<false, false><<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(p);
  B200Q_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200q_backward_qt_bf16(const void* x_e2m1, const void* x_e8m0, const void* rot_bf16, const float* alpha_dev,
                                      void* xh_e2m1, void* xh_e8m0, int size_m, int size_n, int size_b, int flags,
                                      b200q_stream_t stream) {
  int rc = check_device_sm100();
  if (rc) return rc;
  B200Q_REQUIRE(x_e2m1 && x_e8m0 && rot_bf16 && alpha_dev && xh_e2m1 && xh_e8m0, "null pointer argument");
  B200Q_REQUIRE(size_m > 0 && size_n > 0 && size_b > 0, "sizes must be positive");
  B200Q_REQUIRE(size_n % 32 == 0, "size_n (%d) must be a multiple of 32", size_n);
  B200Q_REQUIRE(size_m % 32 == 0, "size_m (%d) must be a multiple of 32", size_m);
  B200Q_REQUIRE(aligned16(x_e2m1) && aligned16(xh_e2m1) && (reinterpret_cast<uintptr_t>(xh_e8m0) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(x_e8m0) & 3) == 0,
                "pointers must be 16-byte aligned (scales: 4-byte)");
  BwdParams p{};
  p.x = x_e2m1; p.x_sf = (const uint8_t*)x_e8m0; p.rot = (const __nv_bfloat16*)rot_bf16; p.alpha = alpha_dev;
  p.q = (uint8_t*)xh_e2m1; p.sf = (uint8_t*)xh_e8m0;
  p.N = size_n; p.M = size_m; p.n_valid = size_n;
  dim3 grid((unsigned)ceil_div(size_n, kBwdTile), (unsigned)ceil_div(size_m, kBwdTile), (unsigned)size_b);
  B200Q_REQUIRE(grid.y <= 65535u && grid.z <= 65535u, "problem too large for one launch");
  // (a generic rotation is bound by its 1024 FMAs per group, not by latency: one-shot form only)
  {
    // B200Q_BWD_QT_TC=0/1 forces the CUDA-core / the tcgen05 kernel (backward_tc.cu: codes decoded to bf16 in shared memory,
    // transpose + rotation on the tensor core)
    const int sw = env().bwd_qt_tc;
    const int64_t numel = (int64_t)size_m * size_n * size_b;
    // Measured (profiles/r02_s3_bwd_bench_final_qt_{cudacore,tensorcore}.jsonl): 43.4 -> 38.8 us at 16384 x 4096, 38.5 -> 35.0 at
    // 4096 x 14336, 13.1 -> 12.4 at 4096 x 4096 -- the decode warps (4 instructions per element pair on ONE warp per scheduler) bound it.
    if ((sw == 1 || (sw < 0 && numel >= ((int64_t)1 << 24))) &&
        backward_qt_tc_eligible(x_e2m1, x_e8m0, rot_bf16, xh_e2m1, xh_e8m0, size_m, size_n, size_b))
      return launch_backward_qt_tc(x_e2m1, x_e8m0, rot_bf16, alpha_dev, xh_e2m1, xh_e8m0, size_m, size_n, size_b, (cudaStream_t)stream);
  }
  if (bwd_pipe_enabled(true, (int64_t)grid.x * grid.y * grid.z) && (flags & B200Q_ROT_TRUSTED_HADAMARD)) return launch_tq_pipe<true, true>(p, grid, (cudaStream_t)stream);
  if (flags & B200Q_ROT_TRUSTED_HADAMARD)
    bwd_transpose_quantize_fp4_kernel<true, true><<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(p);
  else
    bwd_transpose_quantize_fp4_kernel<true, false><<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(p);
  B200Q_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200q_backward_bf16_square_double_mxfp8(const void* x_bf16, int m, int n, void* x_fp8, void* row_scales,
                                                       void* column_scales, b200q_stream_t stream) {
  int rc = check_device_sm100();
  if (rc) return rc;
  B200Q_REQUIRE(x_bf16 && x_fp8 && row_scales && column_scales, "null pointer argument");
  B200Q_REQUIRE(m > 0 && n > 0, "sizes must be positive");
  B200Q_REQUIRE(n % 32 == 0, "n (%d) must be a multiple of 32", n);
  B200Q_REQUIRE(aligned16(x_bf16) && (reinterpret_cast<uintptr_t>(x_fp8) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(row_scales) & 3) == 0 && (reinterpret_cast<uintptr_t>(column_scales) & 3) == 0,
                "pointers must be 16-byte aligned (outputs: 8 / 4-byte)");
  SquareParams p{};
  p.x = (const __nv_bfloat16*)x_bf16; p.y = (uint8_t*)x_fp8; p.row_sf = (uint8_t*)row_scales; p.col_sf = (uint8_t*)column_scales;
  p.m_valid = m; p.m_pad = (int)round_up(m, 128); p.n = n;
  dim3 grid((unsigned)ceil_div(n, 128), (unsigned)(p.m_pad / 128));
  B200Q_REQUIRE(grid.y <= 65535u, "problem too large for one launch");
  bwd_square_double_mxfp8_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(p);
  B200Q_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int b200q_mxfp4_transpose_mxfp8(const void* x_fp4, const void* scales_e8m0, int m, int n, void* x_fp8,
                                           void* shared_exps, b200q_stream_t stream) {
  int rc = check_device_sm100();
  if (rc) return rc;
  B200Q_REQUIRE(x_fp4 && scales_e8m0 && x_fp8 && shared_exps, "null pointer argument");
  B200Q_REQUIRE(m > 0 && n > 0, "sizes must be positive");
  B200Q_REQUIRE(n % 32 == 0, "n (%d) must be a multiple of 32", n);
  B200Q_REQUIRE(aligned16(x_fp4) && aligned16(x_fp8) && (reinterpret_cast<uintptr_t>(shared_exps) & 3) == 0 &&
                    (reinterpret_cast<uintptr_t>(scales_e8m0) & 3) == 0,
                "pointers must be 16-byte aligned (scales: 4-byte)");
  BwdParams p{};
  p.x = x_fp4; p.x_sf = (const uint8_t*)scales_e8m0; p.q = (uint8_t*)x_fp8; p.sf = (uint8_t*)shared_exps;
  p.N = (int)round_up(m, 256); p.M = n; p.n_valid = m;
  dim3 grid((unsigned)(p.N / kBwdTile), (unsigned)ceil_div(n, kBwdTile));
  B200Q_REQUIRE(grid.y <= 65535u, "problem too large for one launch");
  if (bwd_pipe_enabled(true, (int64_t)grid.x * grid.y)) {
    auto kern = bwd_mxfp4_transpose_mxfp8_pipe_kernel;
    static std::atomic<unsigned long long> attr_done{0};
    static std::atomic<int> occ[64];
    if (int rc2 = ensure_dynamic_smem(kern, kBwdTr8PipeSmem, attr_done)) return rc2;
    const int n_tiles = (int)(grid.x * grid.y);
    const int ctas = pipe_grid(kern, kBwdTr8PipeSmem, occ, n_tiles);
    if (ctas <= 0) return B200Q_ECUDA;
    kern<<<ctas, kBwdThreads, kBwdTr8PipeSmem, (cudaStream_t)stream>>>(p, (int)grid.x, n_tiles);
    B200Q_CUDA(cudaGetLastError());
    return 0;
  }
  bwd_mxfp4_transpose_mxfp8_kernel<<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(p);
  B200Q_CUDA(cudaGetLastError());
  return 0;
}
