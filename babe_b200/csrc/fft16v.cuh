// 16-point FFT on float2 values held in registers, forward (e^{-i..}) or inverse (e^{+i..},
// unnormalised) selected at compile time -- no re/im swapping, so the values stay in the 64-bit
// register pairs that Blackwell's two-wide fp32 instructions (FADD2 / FFMA2 / FMUL2) operate on.
// Same decomposition as fft16p (regfft_packed.cuh): 16 = 2 x 8, n = 8a + b, k = c + 2d; natural
// order in and out.  __host__ __device__: tests/host/core4k_host_check.cu runs it on the CPU.
#pragma once
#include "regfft_packed.cuh"

namespace babe {

// v * (-i) for the forward transform, v * (+i) for the inverse
template <bool INV> BABE_HD float2 c_rotq(float2 v) {
  return INV ? make_float2(-v.y, v.x) : make_float2(v.y, -v.x);
}
// a + q b and a - q b with q = -i (forward) / +i (inverse)
template <bool INV> BABE_HD float2 c_add_q(float2 a, float2 b) { return INV ? c_sub_mi(a, b) : c_add_mi(a, b); }
template <bool INV> BABE_HD float2 c_sub_q(float2 a, float2 b) { return INV ? c_add_mi(a, b) : c_sub_mi(a, b); }

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ >= 1000)
// v * (1 + i)/sqrt(2) = ((x - y) h, (x + y) h)
__device__ __forceinline__ float2 c_rot8c(float2 v) {
  constexpr float h = 0.70710678118654752f;
  const float2 t = __ffma2_rn(make_float2(v.y, v.x), make_float2(-1.f, 1.f), v);
  return __fmul2_rn(t, make_float2(h, h));
}
__device__ __forceinline__ float2 c_scale(float2 v, float s) { return __fmul2_rn(v, make_float2(s, s)); }
__device__ __forceinline__ float2 c_fma(float2 v, float s, float2 c) { return __ffma2_rn(v, make_float2(s, s), c); }
__device__ __forceinline__ float2 c_fma2(float2 v, float2 s, float2 c) { return __ffma2_rn(v, s, c); }
#else
__host__ __device__ __forceinline__ float2 c_rot8c(float2 v) {
  constexpr float h = 0.70710678118654752f;
  return make_float2((v.x - v.y) * h, (v.x + v.y) * h);
}
__host__ __device__ __forceinline__ float2 c_scale(float2 v, float s) { return make_float2(v.x * s, v.y * s); }
__host__ __device__ __forceinline__ float2 c_fma(float2 v, float s, float2 c) {
  return make_float2(fmaf(v.x, s, c.x), fmaf(v.y, s, c.y));
}
__host__ __device__ __forceinline__ float2 c_fma2(float2 v, float2 s, float2 c) {
  return make_float2(fmaf(v.x, s.x, c.x), fmaf(v.y, s.y, c.y));
}
#endif
// v * W8^{+-1}
template <bool INV> BABE_HD float2 c_w8(float2 v) { return INV ? c_rot8c(v) : c_rot8(v); }
// Complex products as TWO two-wide instructions (FMUL2 + FFMA2): v w = v w.x + (-v.y, v.x) w.y.  The operand swap,
// the per-half sign and the scalar broadcast are free operand modifiers of FFMA2 (SASS R.F32x2.LO_HI.NP, R.F32),
// and -- measured, profiles/ubench/fp32_rates.cu -- a scalar FFMA issued next to a two-wide one costs the FMA pipe
// about as much as a two-wide one, so mixing 4 scalar instructions per product into the packed butterflies was
// the most expensive way to do it.
// v * w  and  v * conj(w)
#ifndef BABE_CMUL_MODE
#define BABE_CMUL_MODE 2      // bit 0: packed twiddle products, bit 1: packed W16 rotations (all four combinations measure within 1.5 %: the kernels are latency-bound, profiles/r02_apply_filter.md)
#endif
#if BABE_CMUL_MODE & 1
BABE_HD float2 c_mul(float2 v, float2 w) { return c_fma(make_float2(-v.y, v.x), w.y, c_scale(v, w.x)); }
BABE_HD float2 c_mulc(float2 v, float2 w) { return c_fma(make_float2(v.y, -v.x), w.y, c_scale(v, w.x)); }
#else
BABE_HD float2 c_mul(float2 v, float2 w) { return make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x); }
BABE_HD float2 c_mulc(float2 v, float2 w) { return make_float2(v.x * w.x + v.y * w.y, v.y * w.x - v.x * w.y); }
#endif
template <bool INV> BABE_HD float2 c_tw(float2 v, float2 w) { return INV ? c_mulc(v, w) : c_mul(v, w); }
// v * exp(-+ 2 pi i m / 16)
template <bool INV> BABE_HD float2 c_w16(float2 v, int m) {
  const float c = tw_cos16(m), s = tw_sin16(m);
#if BABE_CMUL_MODE & 2
  return INV ? c_fma(make_float2(-v.y, v.x), s, c_scale(v, c)) : c_fma(make_float2(v.y, -v.x), s, c_scale(v, c));
#else
  return INV ? make_float2(v.x * c - v.y * s, v.y * c + v.x * s)
             : make_float2(v.x * c + v.y * s, v.y * c - v.x * s);
#endif
}

template <bool INV> BABE_HD void fft4v(float2& v0, float2& v1, float2& v2, float2& v3) {
  const float2 a0 = c_add(v0, v2), a1 = c_sub(v0, v2);
  const float2 a2 = c_add(v1, v3), a3 = c_sub(v1, v3);
  v0 = c_add(a0, a2);
  v2 = c_sub(a0, a2);
  v1 = c_add_q<INV>(a1, a3);
  v3 = c_sub_q<INV>(a1, a3);
}

constexpr float kH = 0.70710678118654752f;      // 1 / sqrt(2)

// a + s q b and a - s q b (q = -i forward, +i inverse) in ONE two-wide FMA each: the operand swap and the
// sign pattern are free operand modifiers of FFMA2
template <bool INV> BABE_HD float2 c_fma_q(float2 b, float s, float2 a) {
  return INV ? c_fma2(make_float2(b.y, b.x), make_float2(-s, s), a) : c_fma2(make_float2(b.y, b.x), make_float2(s, -s), a);
}
// 4-point transform whose inputs 1 and 3 still carry a pending factor 1/sqrt(2) (v1 = p / sqrt2, v3 = r / sqrt2 are
// never formed: the factor rides on the multiplier of the last FMAs)
template <bool INV> BABE_HD void fft4v_h13(float2& v0, float2& p, float2& v2, float2& r) {
  const float2 a0 = c_add(v0, v2), a1 = c_sub(v0, v2);
  const float2 a2 = c_add(p, r), a3 = c_sub(p, r);
  v0 = c_fma(a2, kH, a0);
  v2 = c_fma(a2, -kH, a0);
  p = c_fma_q<INV>(a3, kH, a1);
  r = c_fma_q<INV>(a3, -kH, a1);
}

// H26: inputs 2 and 6 carry a pending factor 1/sqrt(2)
template <bool INV, bool H26> BABE_HD void fft8v(float2 (&x)[8]) {
  float2 e0 = x[0], e1 = x[2], e2 = x[4], e3 = x[6];
  float2 o0 = x[1], o1 = x[3], o2 = x[5], o3 = x[7];
  if (H26) fft4v_h13<INV>(e0, e1, e2, e3); else fft4v<INV>(e0, e1, e2, e3);
  fft4v<INV>(o0, o1, o2, o3);
  // o1 W8^1 = t1 / sqrt2 and o3 W8^3 = q t3 / sqrt2 with t = o + q o; the 1/sqrt2 goes into the final FMAs
  const float2 t1 = c_add_q<INV>(o1, o1), t3 = c_add_q<INV>(o3, o3);
  x[0] = c_add(e0, o0);          x[4] = c_sub(e0, o0);
  x[1] = c_fma(t1, kH, e1);      x[5] = c_fma(t1, -kH, e1);
  x[2] = c_add_q<INV>(e2, o2);   x[6] = c_sub_q<INV>(e2, o2);   // W8^2 = q
  x[3] = c_fma_q<INV>(t3, kH, e3);
  x[7] = c_fma_q<INV>(t3, -kH, e3);
}

// SCALED: the transform of v[i] * s[i] (window / filter-gain multiply folded into the first butterflies)
template <bool INV, bool SCALED>
BABE_HD void fft16v_impl(float2 (&v)[16], const float (&s)[16]) {
  float2 t0[8], t1[8];
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    float2 d;
    if (SCALED) {
      const float2 lo = c_scale(v[b], s[b]);
      t0[b] = c_fma(v[8 + b], s[8 + b], lo);
      d = c_fma(v[8 + b], -s[8 + b], lo);
    } else {
      t0[b] = c_add(v[b], v[8 + b]);
      d = c_sub(v[b], v[8 + b]);
    }
    if (b == 4) d = c_rotq<INV>(d);                              // W16^4
    else if (b == 2) d = c_add_q<INV>(d, d);                     // W16^2 = W8^1: sqrt2 x the rotated value
    else if (b == 6) { d = c_add_q<INV>(d, d); d = c_rotq<INV>(d); }   // W16^6 = W8^3, likewise
    else if (b != 0) d = c_w16<INV>(d, b);
    t1[b] = d;
  }
  fft8v<INV, false>(t0);
  fft8v<INV, true>(t1);
#pragma unroll
  for (int d = 0; d < 8; ++d) { v[2 * d] = t0[d]; v[2 * d + 1] = t1[d]; }
}
template <bool INV> BABE_HD void fft16v(float2 (&v)[16]) {
  const float none[16] = {};
  fft16v_impl<INV, false>(v, none);
}
template <bool INV> BABE_HD void fft16v_scaled(float2 (&v)[16], const float (&s)[16]) { fft16v_impl<INV, true>(v, s); }

}  // namespace babe
