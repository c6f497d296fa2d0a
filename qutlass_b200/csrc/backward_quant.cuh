// Per-group arithmetic shared by the CUDA-core backward kernels (backward.cu) and the tensor-core one (backward_tc.cu).
#pragma once
#include "quantize_tile.cuh"
#include <cuda_fp16.h>

namespace b200q {

// two e2m1 codes (one byte; element 2i in the low nibble) -> two floats
__device__ __forceinline__ float2 e2m1x2_to_float2(uint32_t byte) {
  uint32_t h2;
  const uint16_t b16 = (uint16_t)byte;
  asm("{\n"
      ".reg .b8 lo8, hi8;\n"
      "mov.b16 {lo8, hi8}, %1;\n"
      "cvt.rn.f16x2.e2m1x2 %0, lo8;\n"
      "}"
      : "=r"(h2)
      : "h"(b16));
  return __half22float2(*reinterpret_cast<const __half2*>(&h2));
}

// eight e2m1 codes (one 32-bit word, element 0 in the low nibble) -> four f16x2 (SASS: F2FP.F16.E2M1.UNPACK_B with the byte
// selectors .B0 ... .B3: no shifts or masks)
__device__ __forceinline__ void e2m1x8_to_half2x4(uint32_t w, uint32_t (&h)[4]) {
  asm("{\n"
      ".reg .b8 b0, b1, b2, b3;\n"
      "mov.b32 {b0, b1, b2, b3}, %4;\n"
      "cvt.rn.f16x2.e2m1x2 %0, b0;\n"
      "cvt.rn.f16x2.e2m1x2 %1, b1;\n"
      "cvt.rn.f16x2.e2m1x2 %2, b2;\n"
      "cvt.rn.f16x2.e2m1x2 %3, b3;\n"
      "}"
      : "=r"(h[0]), "=r"(h[1]), "=r"(h[2]), "=r"(h[3])
      : "r"(w));
}

// two floats -> two e4m3 bytes (RNE, saturate to +-448); `lo` lands in the low byte
__device__ __forceinline__ uint32_t cvt2_e4m3(float lo, float hi) {
  uint16_t r;
  asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(r) : "f"(hi), "f"(lo));
  return r;
}

// shared exponent of the MXFP8 re-quantisers: floor(log2(amax)) - 7 (biased), 127 for an all-zero group
// (quartet_bwd_sm120.cu:497-503 encode_e8m0_shiftm8; tests/quartet_test.py:279-285)
__device__ __forceinline__ uint32_t e8m0_shift7(float amax) {
  return amax == 0.f ? 127u : (((__float_as_uint(amax) >> 23) - 7u) & 0xffu);
}

// exact 2^(127 - e) for a ue8m0 byte e (2^-127 .. 2^127; e = 255 (NaN scale) -> 0)
__device__ __forceinline__ float inv_pow2_of_e8m0(uint32_t e) {
  if (e >= 254u) return e == 254u ? __uint_as_float(0x00400000u) : 0.f;
  return __uint_as_float((254u - e) << 23);
}

// abs-max MXFP4 quantisation of one rotated 32-group: returns the ue8m0 byte, leaves 4 words of packed e2m1 in out[].
//   QT == false (quartet_bwd_sm120.cu:303-315):  s = floor_pow2(amax);          q = e2m1(v * (3 / s))
//   QT == true  (quartet_bwd_sm120.cu:397-410):  s = floor_pow2(amax / alpha);  q = e2m1(v * (3 / (s * alpha)))
// A group whose floored scale is zero (amax == 0 or denormal) yields scale byte 0 and all-zero codes, like the
// reference's test oracle (tests/quartet_test.py:155-175); the reference kernel itself produces NaN -> 0x7 there.
// QT: `c3` = __fdiv_rn(3, alpha), computed once per thread.  The two IEEE divisions per group of the straightforward form
// (amax / alpha for the exponent, 3 / (2^k alpha) for the factor) are reproduced EXACTLY without dividing: the exponent of a
// quotient of two normal numbers is ea - eb - (mantissa_a < mantissa_b) (a quotient in (0.5, 1) never rounds up to 1), and
// 3 / (2^k alpha) = (3 / alpha) 2^-k while nothing leaves the normal range; every other case takes the divisions.
template <bool QT>
__device__ __forceinline__ uint32_t quantise32_absmax(float* v, float alpha, uint32_t* out, float c3 = 0.f) {
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(v[i]));
  uint32_t e;
  float f;
  if constexpr (QT) {
    const uint32_t ua = __float_as_uint(amax), ub = __float_as_uint(alpha), uc = __float_as_uint(c3);
    const int ea = (int)(ua >> 23), eb = (int)((ub >> 23) & 0xffu), ec = (int)((uc >> 23) & 0xffu);
    const int eq = ea - eb + 127 - ((ua & 0x7fffffu) < (ub & 0x7fffffu) ? 1 : 0);
    // normal amax, alpha, 3 / alpha; normal quotient; 2^(eq-127) alpha and (3 / alpha) 2^(127-eq) normal as well
    const bool fast = ea >= 1 && ea <= 254 && eb >= 1 && eb <= 254 && ec >= 1 && ec <= 254 && eq >= 1 && eq <= 253 &&
                      eb + eq - 127 >= 1 && eb + eq - 127 <= 254 && ec + 127 - eq >= 1 && ec + 127 - eq <= 254;
    if (fast) {
      e = (uint32_t)eq;
      f = c3 * __uint_as_float((uint32_t)(254 - eq) << 23);
    } else {
      const float s = __fdiv_rn(amax, alpha);
      e = (__float_as_uint(s) >> 23) & 0xffu;
      const float sp = __uint_as_float(e << 23);
      f = (e == 0u || e == 255u) ? 0.f : __fdiv_rn(3.0f, sp * alpha);
    }
  } else {
    e = (__float_as_uint(amax) >> 23) & 0xffu;
    f = (e == 0u || e == 255u) ? 0.f : 3.0f * inv_pow2_of_e8m0(e);
  }
  const float2 f2 = make_float2(f, f);
#pragma unroll
  for (int i = 0; i < 16; ++i) B200Q_UNPK(v, i, __fmul2_rn(B200Q_PK(v, i), f2));
#pragma unroll
  for (int w = 0; w < 4; ++w) out[w] = cvt8_e2m1(v + 8 * w);
  return e;
}

// cp.async (LDGSTS) with zero fill: `valid == false` copies nothing and writes zeros (the source pointer must still be a valid,
// suitably aligned global address)
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async4_zfill(void* smem_dst, const void* gmem_src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 4 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

}  // namespace b200q
