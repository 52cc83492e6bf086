// 16-point in-register FFT on packed complex values (float2 in one 64-bit
// register pair) using Blackwell's two-wide fp32 instructions (add/fma .f32x2,
// SASS FADD2 / FFMA2 / FMUL2): a complex add or subtract is ONE issue slot
// instead of two.  Same decomposition and twiddles as fft_reg<16> in
// regfft.cuh (16 = 2 x 8), same rounding (every packed op is the two scalar
// round-to-nearest ops side by side), natural order in and out.
#pragma once
#include "regfft.cuh"

namespace babe {

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
__device__ __forceinline__ float2 c_add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 c_sub(float2 a, float2 b) {
  return __ffma2_rn(b, make_float2(-1.f, -1.f), a);
}
// a + (-i) b = (a.x + b.y, a.y - b.x)
__device__ __forceinline__ float2 c_add_mi(float2 a, float2 b) {
  return __ffma2_rn(make_float2(b.y, b.x), make_float2(1.f, -1.f), a);
}
// a - (-i) b = (a.x - b.y, a.y + b.x)
__device__ __forceinline__ float2 c_sub_mi(float2 a, float2 b) {
  return __ffma2_rn(make_float2(b.y, b.x), make_float2(-1.f, 1.f), a);
}
// v * (1 - i)/sqrt(2) = ((x + y) h, (y - x) h)
__device__ __forceinline__ float2 c_rot8(float2 v) {
  constexpr float h = 0.70710678118654752f;
  const float2 t = __ffma2_rn(make_float2(v.y, v.x), make_float2(1.f, -1.f), v);
  return __fmul2_rn(t, make_float2(h, h));
}
#else
__host__ __device__ __forceinline__ float2 c_add(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ float2 c_sub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ float2 c_add_mi(float2 a, float2 b) { return make_float2(a.x + b.y, a.y - b.x); }
__host__ __device__ __forceinline__ float2 c_sub_mi(float2 a, float2 b) { return make_float2(a.x - b.y, a.y + b.x); }
__host__ __device__ __forceinline__ float2 c_rot8(float2 v) {
  constexpr float h = 0.70710678118654752f;
  return make_float2((v.x + v.y) * h, (v.y - v.x) * h);
}
#endif

// v * exp(-2 pi i m / 16), general m (scalar: the operands are single registers anyway)
BABE_HD float2 c_tw16(float2 v, int m) {
  const float c = tw_cos16(m), s = tw_sin16(m);
  return make_float2(v.x * c + v.y * s, v.y * c - v.x * s);
}

BABE_HD void fft4p(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a0 = c_add(v0, v2), a1 = c_sub(v0, v2);
  const float2 a2 = c_add(v1, v3), a3 = c_sub(v1, v3);
  v0 = c_add(a0, a2);
  v2 = c_sub(a0, a2);
  v1 = c_add_mi(a1, a3);
  v3 = c_sub_mi(a1, a3);
}

// natural order in (x0..x7) and out
BABE_HD void fft8p(float2 (&x)[8]) {
  float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
  float2 o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
  fft4p(e0, e1, e2, e3);
  fft4p(o0, o1, o2, o3);
  o1 = c_rot8(o1);            // W8^1
  o3 = c_rot8(o3);            // W8^3 = W8^1 * (-i): the -i goes into the final add
  x[0] = c_add(e0, o0);    x[4] = c_sub(e0, o0);
  x[1] = c_add(e1, o1);    x[5] = c_sub(e1, o1);
  x[2] = c_add_mi(e2, o2); x[6] = c_sub_mi(e2, o2);   // W8^2 = -i
  x[3] = c_add_mi(e3, o3); x[7] = c_sub_mi(e3, o3);
}

// 16 = 2 x 8: n = 8a + b, k = c + 2d
BABE_HD void fft16p(float2 (&v)[16]) {
  float2 t0[8], t1[8];          // c = 0 and c = 1 after the 2-point stage
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    t0[b] = c_add(v[b], v[8 + b]);
    float2 d = c_sub(v[b], v[8 + b]);
    if (b == 4) d = make_float2(d.y, -d.x);            // W16^4 = -i
    else if (b == 2) d = c_rot8(d);                     // W16^2 = W8^1
    else if (b == 6) { d = c_rot8(d); d = make_float2(d.y, -d.x); }   // W16^6 = W8^3
    else if (b != 0) d = c_tw16(d, b);
    t1[b] = d;
  }
  fft8p(t0);
  fft8p(t1);
#pragma unroll
  for (int d = 0; d < 8; ++d) { v[2 * d] = t0[d]; v[2 * d + 1] = t1[d]; }
}

// split-array front ends matching fft_reg<16>(re, im); the inverse is obtained
// by passing (im, re)
BABE_HD void fft16_split(float (&re)[16], float (&im)[16]) {
  float2 v[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = make_float2(re[i], im[i]);
  fft16p(v);
#pragma unroll
  for (int i = 0; i < 16; ++i) { re[i] = v[i].x; im[i] = v[i].y; }
}

}  // namespace babe
