// In-register complex FFTs of 2..64 points (forward, e^{-i...}; natural in,
// natural out).  Every index is a compile-time constant after unrolling, so
// the arrays live in registers and the twiddles fold into immediates.
// The inverse transform is obtained by swapping the re/im arrays on the way
// in and out (IFFT(z) = swap(FFT(swap(z)))), which costs nothing.
#pragma once
#include "fft_consts.cuh"

#ifndef BABE_HD
#define BABE_HD __host__ __device__ __forceinline__
#endif

namespace babe {

template <int R> BABE_HD constexpr float tw_cos(int m) {
  return R == 16 ? tw_cos16(m % 16) : R == 32 ? tw_cos32(m % 32) : tw_cos64(m % 64);
}
template <int R> BABE_HD constexpr float tw_sin(int m) {
  return R == 16 ? tw_sin16(m % 16) : R == 32 ? tw_sin32(m % 32) : tw_sin64(m % 64);
}

// (r,i) *= exp(-2*pi*i*m/R), with the trivial rotations special-cased
template <int R> BABE_HD void rot(float& r, float& i, int m) {
  m %= R;
  if (m == 0) return;
  if (4 * m == R) { float t = r; r = i; i = -t; return; }           // * (-i)
  if (2 * m == R) { r = -r; i = -i; return; }                       // * (-1)
  if (4 * m == 3 * R) { float t = r; r = -i; i = t; return; }       // * (+i)
  const float c = tw_cos<R>(m), s = tw_sin<R>(m);
  if (8 * m == R) { const float t = (r + i) * c; i = (i - r) * c; r = t; return; }
  const float nr = r * c + i * s;
  const float ni = i * c - r * s;
  r = nr; i = ni;
}

BABE_HD void fft2(float (&r)[2], float (&i)[2]) {
  const float tr = r[0] - r[1], ti = i[0] - i[1];
  r[0] += r[1]; i[0] += i[1]; r[1] = tr; i[1] = ti;
}

BABE_HD void fft4(float (&r)[4], float (&i)[4]) {
  const float a0r = r[0] + r[2], a0i = i[0] + i[2];
  const float a1r = r[0] - r[2], a1i = i[0] - i[2];
  const float a2r = r[1] + r[3], a2i = i[1] + i[3];
  const float a3r = r[1] - r[3], a3i = i[1] - i[3];
  r[0] = a0r + a2r; i[0] = a0i + a2i;
  r[2] = a0r - a2r; i[2] = a0i - a2i;
  r[1] = a1r + a3i; i[1] = a1i - a3r;
  r[3] = a1r - a3i; i[3] = a1i + a3r;
}

BABE_HD void fft8(float (&r)[8], float (&i)[8]) {
  float er[4] = {r[0], r[2], r[4], r[6]}, ei[4] = {i[0], i[2], i[4], i[6]};
  float od[4] = {r[1], r[3], r[5], r[7]}, oi[4] = {i[1], i[3], i[5], i[7]};
  fft4(er, ei);
  fft4(od, oi);
  constexpr float h = 0.70710678118654752f;
  // W8^1 = (1-i)/sqrt2, W8^2 = -i, W8^3 = (-1-i)/sqrt2
  { const float t = (od[1] + oi[1]) * h; oi[1] = (oi[1] - od[1]) * h; od[1] = t; }
  { const float t = od[2]; od[2] = oi[2]; oi[2] = -t; }
  { const float t = (oi[3] - od[3]) * h; oi[3] = -(od[3] + oi[3]) * h; od[3] = t; }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    r[k] = er[k] + od[k]; i[k] = ei[k] + oi[k];
    r[k + 4] = er[k] - od[k]; i[k + 4] = ei[k] - oi[k];
  }
}

template <int P> BABE_HD void fft_small(float (&r)[P], float (&i)[P]);
template <> BABE_HD void fft_small<2>(float (&r)[2], float (&i)[2]) { fft2(r, i); }
template <> BABE_HD void fft_small<4>(float (&r)[4], float (&i)[4]) { fft4(r, i); }
template <> BABE_HD void fft_small<8>(float (&r)[8], float (&i)[8]) { fft8(r, i); }

// R = P*8 point FFT, R in {16, 32, 64}.
template <int R> BABE_HD void fft_reg(float (&re)[R], float (&im)[R]) {
  constexpr int Q = 8, P = R / Q;
  // step 1: P-point FFTs over a (n = Q*a + b), twiddle W_R^{b c}, keep t[b][c] at Q*c + b
#pragma unroll
  for (int b = 0; b < Q; ++b) {
    float tr[P], ti[P];
#pragma unroll
    for (int a = 0; a < P; ++a) { tr[a] = re[Q * a + b]; ti[a] = im[Q * a + b]; }
    fft_small<P>(tr, ti);
#pragma unroll
    for (int c = 0; c < P; ++c) {
      rot<R>(tr[c], ti[c], b * c);
      re[Q * c + b] = tr[c]; im[Q * c + b] = ti[c];
    }
  }
  // step 2: Q-point FFTs over b for each c; X[c + P d] lands at Q*c + d
  float outr[R], outi[R];
#pragma unroll
  for (int c = 0; c < P; ++c) {
    float tr[Q], ti[Q];
#pragma unroll
    for (int b = 0; b < Q; ++b) { tr[b] = re[Q * c + b]; ti[b] = im[Q * c + b]; }
    fft8(tr, ti);
#pragma unroll
    for (int d = 0; d < Q; ++d) { outr[c + P * d] = tr[d]; outi[c + P * d] = ti[d]; }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) { re[k] = outr[k]; im[k] = outi[k]; }
}

}  // namespace babe
