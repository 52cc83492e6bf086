// Mixed-radix Stockham FFT over sequences held in shared memory.
//
// nseq independent complex sequences of length n (row stride `stride`, odd in
// float2 units to stay bank-conflict free) are transformed by all threads of
// the CTA.  Radices 16/8/4/2 use the in-register butterflies of regfft.cuh;
// odd primes up to 23 use a direct r x r DFT with the r-th roots taken from
// the n-th root table (n is a multiple of r), so any length whose prime
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

struct FftFactors {
  int n;
  int nf;
  int radix[MAX_FACTORS];
};

BABE_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
BABE_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

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

template <int R> BABE_HD void butterfly(float (&vr)[R], float (&vi)[R], const float2* wn, int n) {
  dft_direct<R>(vr, vi, wn, n);
}
template <> BABE_HD void butterfly<2>(float (&vr)[2], float (&vi)[2], const float2*, int) { fft2(vr, vi); }
template <> BABE_HD void butterfly<4>(float (&vr)[4], float (&vi)[4], const float2*, int) { fft4(vr, vi); }
template <> BABE_HD void butterfly<8>(float (&vr)[8], float (&vi)[8], const float2*, int) { fft8(vr, vi); }
template <> BABE_HD void butterfly<16>(float (&vr)[16], float (&vi)[16], const float2*, int) { fft_reg<16>(vr, vi); }

// One Stockham stage: radix R, Ns = product of the radices already applied.
template <int R>
BABE_HD void stockham_stage(const float2* in, float2* out, int n, int stride, int nseq, int Ns,
                            const float2* wn, int tid, int nthreads) {
  const int m = n / R;
  const int tw_step = n / (Ns * R);
  const int tasks = nseq * m;
  for (int q = tid; q < tasks; q += nthreads) {
    const int seq = q / m, j = q - seq * m;
    const int k = j % Ns;
    const float2* src = in + seq * stride + j;
    float vr[R], vi[R];
#pragma unroll
    for (int t = 0; t < R; ++t) {
      float2 v = src[t * m];
      if (t > 0 && k > 0) v = cmul(v, wn[t * k * tw_step]);
      vr[t] = v.x; vi[t] = v.y;
    }
    butterfly<R>(vr, vi, wn, n);
    float2* dst = out + seq * stride + (j - k) * R + k;
#pragma unroll
    for (int t = 0; t < R; ++t) dst[t * Ns] = make_float2(vr[t], vi[t]);
  }
}

// Odd-prime stage with one OUTPUT per task (R x more parallelism than one
// butterfly per task, which leaves most of the CTA idle for R = 13..23):
//   out[u] = sum_t in[t] W_n^{t (k step + u n/R)}
// -- the stage twiddle and the DFT weight are one root lookup per term.
template <int R>
BABE_HD void stockham_stage_wide(const float2* in, float2* out, int n, int stride, int nseq,
                                 int Ns, const float2* wn, int tid, int nthreads) {
  const int m = n / R;
  const int tw_step = n / (Ns * R);
  const int nr = n / R;
  const int tasks = nseq * n;
  for (int q = tid; q < tasks; q += nthreads) {
    const int seq = q / n, rem = q - seq * n;
    const int u = rem / m, j = rem - u * m;          // consecutive threads: consecutive j
    const int k = j % Ns;
    int inc = k * tw_step + u * nr;
    if (inc >= n) inc -= n;
    const float2* src = in + seq * stride + j;
    // all 2(R-1) shared-memory loads are issued before the multiply-adds (ILP)
    float2 v[R], w[R];
    v[0] = src[0];
    int idx = 0;
#pragma unroll
    for (int t = 1; t < R; ++t) {
      idx += inc;
      if (idx >= n) idx -= n;
      v[t] = src[t * m];
      w[t] = wn[idx];
    }
    float2 acc = v[0], acc2 = make_float2(0.f, 0.f);
#pragma unroll
    for (int t = 1; t < R; ++t) {
      float2& a = (t & 1) ? acc : acc2;               // two accumulation chains
      a.x += v[t].x * w[t].x - v[t].y * w[t].y;
      a.y += v[t].x * w[t].y + v[t].y * w[t].x;
    }
    out[seq * stride + (j - k) * R + k + u * Ns] = make_float2(acc.x + acc2.x, acc.y + acc2.y);
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
      case 2: stockham_stage<2>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 3: stockham_stage<3>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 4: stockham_stage<4>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 5: stockham_stage<5>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 7: stockham_stage_wide<7>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 8: stockham_stage<8>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 11: stockham_stage_wide<11>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 13: stockham_stage_wide<13>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 16: stockham_stage<16>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 17: stockham_stage_wide<17>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 19: stockham_stage_wide<19>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      case 23: stockham_stage_wide<23>(src, dst, f.n, stride, nseq, Ns, wn, tid, nthreads); break;
      default: break;
    }
    Ns *= r;
    BABE_CTA_SYNC();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

}  // namespace babe
