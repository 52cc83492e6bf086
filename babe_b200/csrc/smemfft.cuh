// Mixed-radix Stockham FFT over sequences held in shared memory.
//
// nseq independent complex sequences of length n (row stride `stride`, odd in
// float2 units to stay bank-conflict free) are transformed by all threads of
// the CTA.  Radices 16/8/4/2 use the in-register butterflies of regfft.cuh;
// odd primes up to 23 use a conjugate-symmetric small DFT with compile-time
// roots (dft_odd_sym), so any length whose prime
// factors are <= 23 is supported -- enough for the non-power-of-two segment
// lengths of the reference (184184 = 2^3*7*11*13*23, 132300 = 2^2*3^3*5^2*7^2,
// 368368, 485100; conf/exp/*.yaml).
//
// The routines are __host__ __device__ so that the arithmetic can be verified
// on the CPU; `tid`/`nthreads` are threadIdx.x/blockDim.x on the device.
#pragma once
#include "regfft.cuh"

namespace babe {

constexpr int MAX_FACTORS = 12;

// Division by a small runtime constant as one wide multiply: q / d = (q * M) >> 40 with
// M = ceil(2^40 / d), exact for q < 2^20 and d < 2^12 (all task counts here are far below).
struct FastDiv {
  unsigned long long M;
  int d;
  BABE_HD int div(int q) const { return (int)(((unsigned long long)(unsigned)q * M) >> 40); }
  BABE_HD int mod(int q) const { return q - div(q) * d; }
};
inline FastDiv make_fastdiv(int d) {
  FastDiv f;
  f.d = d;
  f.M = ((1ull << 40) + (unsigned long long)d - 1) / (unsigned long long)d;
  return f;
}

struct FftFactors {
  int n;
  int nf;
  int radix[MAX_FACTORS];
  FastDiv div_m[MAX_FACTORS];    // by n / radix[s]
  FastDiv div_ns[MAX_FACTORS];   // by the product of the radices before stage s
};
inline void fill_fastdiv(FftFactors& f) {
  int ns = 1;
  for (int s = 0; s < f.nf; ++s) {
    f.div_m[s] = make_fastdiv(f.n / f.radix[s]);
    f.div_ns[s] = make_fastdiv(ns);
    ns *= f.radix[s];
  }
}

BABE_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
BABE_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// Sequence element a lives at physical slot a + a/16: the first Stockham stage of
// radix R writes with stride R (16 float2 = 128 B = all 32 banks), which without
// the skew is a 16..32-way bank conflict (66 % of all wavefronts in the round-1
// profile of k_cqt_analysis).
BABE_HD int pad16(int a) { return a + (a >> 4); }
BABE_HD int padded_len(int n) { return (n + (n >> 4) + 1) | 1; }

template <int R>
BABE_HD void dft_direct(float (&vr)[R], float (&vi)[R], const float2* wn, int n) {
  // r-th roots W_r^m = W_n^{m n/r}
  float wr[R], wi[R];
  const int step = n / R;
#pragma unroll
  for (int m = 0; m < R; ++m) { const float2 w = wn[m * step]; wr[m] = w.x; wi[m] = w.y; }
  float orr[R], oi[R];
#pragma unroll
  for (int u = 0; u < R; ++u) {
    float sr = vr[0], si = vi[0];
#pragma unroll
    for (int t = 1; t < R; ++t) {
      const int m = (t * u) % R;
      sr += vr[t] * wr[m] - vi[t] * wi[m];
      si += vr[t] * wi[m] + vi[t] * wr[m];
    }
    orr[u] = sr; oi[u] = si;
  }
#pragma unroll
  for (int u = 0; u < R; ++u) { vr[u] = orr[u]; vi[u] = oi[u]; }
}

// Odd-prime DFT exploiting W^{(R-t)u} = conj(W^{tu}): with a_t = v_t + v_{R-t},
// b_t = v_t - v_{R-t},  X_u = v_0 + sum_t a_t cos(2 pi t u/R) - i sum_t b_t sin(2 pi t u/R)
// and X_{R-u} is the same with +i.  (R-1)^2 + O(R) real multiply-adds instead of
// 4 R^2, and the roots are compile-time immediates (no table, no registers).
template <int R>
BABE_HD void dft_odd_sym(float (&vr)[R], float (&vi)[R]) {
  constexpr int H = (R - 1) / 2;
  float ar[H], ai[H], br[H], bi[H];
#pragma unroll
  for (int t = 1; t <= H; ++t) {
    ar[t - 1] = vr[t] + vr[R - t]; ai[t - 1] = vi[t] + vi[R - t];
    br[t - 1] = vr[t] - vr[R - t]; bi[t - 1] = vi[t] - vi[R - t];
  }
  const float x0r = vr[0], x0i = vi[0];
  float s0r = x0r, s0i = x0i;
#pragma unroll
  for (int t = 0; t < H; ++t) { s0r += ar[t]; s0i += ai[t]; }
  vr[0] = s0r; vi[0] = s0i;
#pragma unroll
  for (int u = 1; u <= H; ++u) {
    float Ar = x0r, Ai = x0i, Br = 0.f, Bi = 0.f;
#pragma unroll
    for (int t = 1; t <= H; ++t) {
      const int m = (t * u) % R;
      const float c = odd_cos<R>(m), sn = odd_sin<R>(m);
      Ar += ar[t - 1] * c; Ai += ai[t - 1] * c;
      Br += br[t - 1] * sn; Bi += bi[t - 1] * sn;
    }
    // X_u = A - i B,  X_{R-u} = A + i B
    vr[u] = Ar + Bi; vi[u] = Ai - Br;
    vr[R - u] = Ar - Bi; vi[R - u] = Ai + Br;
  }
}

template <int R> BABE_HD void butterfly(float (&vr)[R], float (&vi)[R], const float2* wn, int n) {
  dft_odd_sym<R>(vr, vi);
}
template <> BABE_HD void butterfly<2>(float (&vr)[2], float (&vi)[2], const float2*, int) { fft2(vr, vi); }
template <> BABE_HD void butterfly<4>(float (&vr)[4], float (&vi)[4], const float2*, int) { fft4(vr, vi); }
template <> BABE_HD void butterfly<8>(float (&vr)[8], float (&vi)[8], const float2*, int) { fft8(vr, vi); }
template <> BABE_HD void butterfly<16>(float (&vr)[16], float (&vi)[16], const float2*, int) { fft_reg<16>(vr, vi); }

// One Stockham stage: radix R, Ns = product of the radices already applied.
template <int R>
BABE_HD void stockham_stage(const float2* in, float2* out, int n, int stride, int nseq, int Ns,
                            const float2* wn, int tid, int nthreads, FastDiv dm, FastDiv dns) {
  const int m = dm.d;
  const int tw_step = m / Ns;
  const int tasks = nseq * m;
  for (int q = tid; q < tasks; q += nthreads) {
    const int seq = dm.div(q), j = q - seq * m;
    const int k = dns.mod(j);
    const float2* src = in + seq * stride;
    float vr[R], vi[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      float2 v = src[pad16(j + t * m)];
      if (t > 0 && k > 0) v = cmul(v, wn[t * k * tw_step]);
      vr[t] = v.x; vi[t] = v.y;
    }
    butterfly<R>(vr, vi, wn, n);
    float2* dst = out + seq * stride;
    const int o0 = (j - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) dst[pad16(o0 + t * Ns)] = make_float2(vr[t], vi[t]);
  }
}

#ifdef __CUDA_ARCH__
#define BABE_CTA_SYNC() __syncthreads()
#else
#define BABE_CTA_SYNC() ((void)0)
#endif

// Forward FFT of nseq sequences; data starts in `a`, returns the buffer that
// holds the result (a or b).  Ends with a CTA barrier.
BABE_HD float2* smem_fft(float2* a, float2* b, const FftFactors& f, int stride, int nseq,
                         const float2* wn, int tid, int nthreads) {
  int Ns = 1;
  float2* src = a;
  float2* dst = b;
  for (int s = 0; s < f.nf; ++s) {
    const int r = f.radix[s];
    switch (r) {
      case 2: stockham_stage<2>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 3: stockham_stage<3>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 4: stockham_stage<4>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 5: stockham_stage<5>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 7: stockham_stage<7>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 8: stockham_stage<8>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 11: stockham_stage<11>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 13: stockham_stage<13>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 16: stockham_stage<16>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 17: stockham_stage<17>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 19: stockham_stage<19>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      case 23: stockham_stage<23>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads, f.div_m[s], f.div_ns[s]); break;
      default: break;
    }
    Ns *= r;
    BABE_CTA_SYNC();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

}  // namespace babe
