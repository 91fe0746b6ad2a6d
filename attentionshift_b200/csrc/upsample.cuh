// Bilinear resampling with the exact arithmetic of torch's CPU kernel (align_corners=False), evaluated on the fly.
//   src = max(scale*(dst+0.5)-0.5, 0); i0 = int(src); i1 = min(i0+1, n-1); w1 = src - i0; w0 = 1 - w1
//   t   = fma(v[i0], wx0, v[i1]*wx1)   (inner dim);   out = fma(t_y0, wy0, t_y1*wy1)
// (the association GCC picks for `t0*w0 + t1*w1` in ATen/native/cpu/UpSampleKernel.cpp; checked bit-for-bit against
//  F.interpolate on CPU in tests/test_upsample_formula.py).
#pragma once
#include <cuda_runtime.h>

namespace asb {

struct Tap { int i0, i1; float w0, w1; };

// x16 up-sampling tap for destination index d of a source axis with n samples
__device__ __forceinline__ Tap tap_up16(int d, int n) {
  const float src = fmaxf(__fsub_rn(__fmul_rn(0.0625f, (float)d + 0.5f), 0.5f), 0.f);
  Tap t;
  t.i0 = (int)src;
  t.i1 = min(t.i0 + 1, n - 1);
  t.w1 = __fsub_rn(src, (float)t.i0);
  t.w0 = __fsub_rn(1.f, t.w1);
  return t;
}

__device__ __forceinline__ float lerp2(float a, float b, float c, float d, const Tap& ty, const Tap& tx) {
  const float t0 = __fmaf_rn(a, tx.w0, __fmul_rn(b, tx.w1));
  const float t1 = __fmaf_rn(c, tx.w0, __fmul_rn(d, tx.w1));
  return __fmaf_rn(t0, ty.w0, __fmul_rn(t1, ty.w1));
}

// value at pixel (y, x) of the x16 up-sampling of low[hp][wp]
__device__ __forceinline__ float up16(const float* low, int hp, int wp, int y, int x) {
  const Tap ty = tap_up16(y, hp), tx = tap_up16(x, wp);
  const float* r0 = low + ty.i0 * wp;
  const float* r1 = low + ty.i1 * wp;
  return lerp2(r0[tx.i0], r0[tx.i1], r1[tx.i0], r1[tx.i1], ty, tx);
}

}  // namespace asb
