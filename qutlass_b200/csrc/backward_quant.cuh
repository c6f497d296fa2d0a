// Per-group arithmetic shared by the CUDA-core backward kernels (backward.cu) and the tensor-core one (backward_tc.cu).
#pragma once
#include "quantize_tile.cuh"

namespace b200q {

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
template <bool QT>
__device__ __forceinline__ uint32_t quantise32_absmax(float* v, float alpha, uint32_t* out) {
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) amax = fmaxf(amax, fabsf(v[i]));
  float s = amax;
  if constexpr (QT) s = __fdiv_rn(amax, alpha);
  const uint32_t e = (__float_as_uint(s) >> 23) & 0xffu;
  float f;
  if constexpr (QT) {
    const float sp = __uint_as_float(e << 23);
    f = (e == 0u || e == 255u) ? 0.f : __fdiv_rn(3.0f, sp * alpha);
  } else {
    f = (e == 0u || e == 255u) ? 0.f : 3.0f * inv_pow2_of_e8m0(e);
  }
  const float2 f2 = make_float2(f, f);
#pragma unroll
  for (int i = 0; i < 16; ++i) B200Q_UNPK(v, i, __fmul2_rn(B200Q_PK(v, i), f2));
#pragma unroll
  for (int w = 0; w < 4; ++w) out[w] = cvt8_e2m1(v + 8 * w);
  return e;
}

}  // namespace b200q
