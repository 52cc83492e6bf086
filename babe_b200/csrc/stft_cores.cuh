// Transform cores of the STFT-domain kernels (stft_ops.cu): N-point complex FFT / inverse FFT of one
// frame pair held in the registers of a thread group, with shared-memory exchanges between the
// radix passes.  Core3's passes are __host__ __device__ so that tests/host/core3_host_check.cu can
// emulate the 256 threads of a group on the CPU.
#pragma once
#include "common.cuh"
#include "regfft.cuh"
#include "regfft_packed.cuh"

namespace babe {

BABE_HD float2 cmulf(float ar, float ai, float2 w) {
  return make_float2(ar * w.x - ai * w.y, ar * w.y + ai * w.x);
}
BABE_HD float2 cmulcf(float ar, float ai, float2 w) {   // * conj(w)
  return make_float2(ar * w.x + ai * w.y, ai * w.x - ar * w.y);
}

// ---------------------------------------------------------------------------
// Core interface (all static):
//   N, HOP, F, TPF                      threads per frame group
//   NT, TS, TT   time role : thread t < TT holds z[TS*i + t], i < NT
//   NF, KS, FT   freq role : thread t < FT holds Z[t + KS*i], i < NF
//   EX_ELEMS, TW_SMEM                   float2 elements of exchange / twiddle smem
//   Regs, init_regs()                   per-thread twiddle constants
//   load_twiddles()                     fill the shared twiddle table (CTA-wide)
//   fwd(), inv(), mirror()
// `roots` is the table exp(-2 pi i m / N), m < N.
// Callers must pass a group barrier between two uses of `ex` (fwd/inv/mirror
// each end with reads of ex by other threads' data).
// ---------------------------------------------------------------------------
template <int R1_, int R2_>
struct Core2 {
  static constexpr int R1 = R1_, R2 = R2_;
  static constexpr int N = R1 * R2, HOP = N / 2, F = N / 2 + 1;
  static constexpr int TPF = R1 > R2 ? R1 : R2;
  static constexpr int MIN_CTAS = 1;                         // register-heavy: one CTA per SM
  static constexpr int NT = R1, TS = R2, TT = R2;
  static constexpr int NF = R2, KS = R1, FT = R1;
  static constexpr int EXF = R2 + 1;                        // forward exchange  [k1][n2]
  static constexpr int EXI = R1 + 1;                        // inverse exchange  [n2][k1]
  static constexpr int EX_ELEMS = (R1 * EXF > R2 * EXI) ? R1 * EXF : R2 * EXI;
  static constexpr int TW_SMEM = R1 * EXF;                  // padded twiddle table [k1][n2]
  struct Regs {};
  __device__ static __forceinline__ void init_regs(Regs&, const float2*, int) {}
  __device__ static __forceinline__ void load_twiddles(float2* tw, const float2* roots) {
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
      const int k1 = i / R2, n2 = i % R2;
      tw[k1 * EXF + n2] = roots[k1 * n2];
    }
  }
  __device__ static __forceinline__ void fwd(float (&ar)[NT], float (&ai)[NT], float (&br)[NF],
                                             float (&bi)[NF], float2* ex, const float2* tw,
                                             const Regs&, int t, int bar) {
    if (t < TT) {
      fft_reg<R1>(ar, ai);
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) ex[k1 * EXF + t] = cmulf(ar[k1], ai[k1], tw[k1 * EXF + t]);
    }
    group_sync<TPF>(bar);
    if (t < FT) {
#pragma unroll
      for (int n2 = 0; n2 < R2; ++n2) {
        const float2 v = ex[t * EXF + n2];
        br[n2] = v.x; bi[n2] = v.y;
      }
      fft_reg<R2>(br, bi);
    }
  }
  // in : Z[t + R1*k2]; out: N * z[R2*n1 + t] (unnormalised)
  __device__ static __forceinline__ void inv(float (&br)[NF], float (&bi)[NF], float (&ar)[NT],
                                             float (&ai)[NT], float2* ex, const float2* tw,
                                             const Regs&, int t, int bar) {
    if (t < FT) {
      fft_reg<R2>(bi, br);
#pragma unroll
      for (int n2 = 0; n2 < R2; ++n2) ex[n2 * EXI + t] = cmulcf(br[n2], bi[n2], tw[t * EXF + n2]);
    }
    group_sync<TPF>(bar);
    if (t < TT) {
#pragma unroll
      for (int k1 = 0; k1 < R1; ++k1) {
        const float2 v = ex[t * EXI + k1];
        ar[k1] = v.x; ai[k1] = v.y;
      }
      fft_reg<R1>(ai, ar);
    }
  }
  // (pr,pi)[k2] <- Z[N - (t + R1*k2)]
  __device__ static __forceinline__ void mirror(const float (&br)[NF], const float (&bi)[NF],
                                                float (&pr)[NF], float (&pi)[NF], float2* ex, int t,
                                                int bar) {
    if (t < FT) {
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) ex[t * EXF + k2] = make_float2(br[k2], bi[k2]);
    }
    group_sync<TPF>(bar);
    if (t < FT) {
      const int pt = (R1 - t) % R1;
#pragma unroll
      for (int k2 = 0; k2 < R2; ++k2) {
        const int pk = (t == 0) ? ((R2 - k2) % R2) : (R2 - 1 - k2);
        const float2 v = ex[pt * EXF + pk];
        pr[k2] = v.x; pi[k2] = v.y;
      }
    }
  }
};

// N = 4096 = 16 * 16 * 16.  n = 256 n1 + 16 n2 + n3, k = k1 + 16 k2 + 256 k3.
//   P1: thread (n2,n3) = t         FFT over n1 -> k1, * W_4096^{t k1}   -> ex[k1][t]
//   P2: thread (k1 = t/16, n3)     FFT over n2 -> k2, in place in ex
//   P3: thread (k1 = t%16, k2)     * W_256^{n3 k2}, FFT over n3 -> k3   => Z[t + 256 k3]
// ex rows are padded to 257 float2 so that all three access patterns are
// bank-conflict free.  The inverse runs the same passes backwards.
struct Core3 {
  static constexpr int N = 4096, HOP = 2048, F = 2049;
  static constexpr int TPF = 256;
  static constexpr int MIN_CTAS = 2;                         // <= 128 registers: two CTAs per SM
  static constexpr int NT = 16, TS = 256, TT = 256;
  static constexpr int NF = 16, KS = 256, FT = 256;
  static constexpr int ROW = 257;
  static constexpr int EX_ELEMS = 16 * ROW;
  static constexpr int TW_SMEM = 256;                       // W_256^m
  struct Regs { float wr[16], wi[16]; };                    // W_4096^{t k1}
  BABE_HD static void init_regs(Regs& r, const float2* roots, int t) {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float2 w = roots[t * k1];
      r.wr[k1] = w.x; r.wi[k1] = w.y;
    }
  }
  __device__ static __forceinline__ void load_twiddles(float2* tw, const float2* roots) {
    for (int i = threadIdx.x; i < 256; i += blockDim.x) tw[i] = roots[16 * i];
  }
  // The passes as per-thread functions; between two of them every thread of the group must have
  // finished the previous one (group barrier on the device; the host test in
  // tests/host/core3_host_check.cu runs the 256 "threads" one after the other).
  BABE_HD static void fwd_p1(float (&ar)[16], float (&ai)[16], float2* ex, const Regs& rg, int t) {
    fft16_split(ar, ai);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1)
      ex[k1 * ROW + t] = make_float2(ar[k1] * rg.wr[k1] - ai[k1] * rg.wi[k1],
                                     ar[k1] * rg.wi[k1] + ai[k1] * rg.wr[k1]);
  }
  BABE_HD static void fwd_p2(float (&br)[16], float (&bi)[16], float2* ex, int t) {
    float2* col = ex + (t >> 4) * ROW + (t & 15);          // + 16 n2
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) { const float2 v = col[16 * n2]; br[n2] = v.x; bi[n2] = v.y; }
    fft16_split(br, bi);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) col[16 * k2] = make_float2(br[k2], bi[k2]);
  }
  BABE_HD static void fwd_p3(float (&br)[16], float (&bi)[16], const float2* ex, const float2* tw, int t) {
    const int k2 = t >> 4;
    const float2* row = ex + (t & 15) * ROW + 16 * k2;      // + n3
#pragma unroll
    for (int n3 = 0; n3 < 16; ++n3) {
      const float2 v = cmulf(row[n3].x, row[n3].y, tw[n3 * k2]);
      br[n3] = v.x; bi[n3] = v.y;
    }
    fft16_split(br, bi);
  }
  BABE_HD static void inv_p1(float (&br)[16], float (&bi)[16], float2* ex, const float2* tw, int t) {
    const int k2 = t >> 4;
    float2* row = ex + (t & 15) * ROW + 16 * k2;
    fft16_split(bi, br);                                    // inverse over k3 -> n3
#pragma unroll
    for (int n3 = 0; n3 < 16; ++n3) row[n3] = cmulcf(br[n3], bi[n3], tw[n3 * k2]);
  }
  BABE_HD static void inv_p2(float (&br)[16], float (&bi)[16], float2* ex, int t) {
    float2* col = ex + (t >> 4) * ROW + (t & 15);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) { const float2 v = col[16 * k2]; br[k2] = v.x; bi[k2] = v.y; }
    fft16_split(bi, br);                                    // inverse over k2 -> n2
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) col[16 * n2] = make_float2(br[n2], bi[n2]);
  }
  BABE_HD static void inv_p3(float (&ar)[16], float (&ai)[16], const float2* ex, const Regs& rg, int t) {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float2 v = ex[k1 * ROW + t];
      ar[k1] = v.x * rg.wr[k1] + v.y * rg.wi[k1];             // * conj(W^{t k1})
      ai[k1] = v.y * rg.wr[k1] - v.x * rg.wi[k1];
    }
    fft16_split(ai, ar);                                      // inverse over k1 -> n1
  }
  BABE_HD static void mirror_store(const float (&br)[16], const float (&bi)[16], float2* ex, int t) {
#pragma unroll
    for (int i = 0; i < 16; ++i) ex[i * ROW + t] = make_float2(br[i], bi[i]);
  }
  BABE_HD static void mirror_load(float (&pr)[16], float (&pi)[16], const float2* ex, int t) {
    const int pt = (256 - t) & 255;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int pk = (t == 0) ? ((16 - i) & 15) : (15 - i);
      const float2 v = ex[pk * ROW + pt];
      pr[i] = v.x; pi[i] = v.y;
    }
  }
  __device__ static __forceinline__ void fwd(float (&ar)[16], float (&ai)[16], float (&br)[16],
                                             float (&bi)[16], float2* ex, const float2* tw,
                                             const Regs& rg, int t, int bar) {
    fwd_p1(ar, ai, ex, rg, t);
    group_sync<TPF>(bar);
    fwd_p2(br, bi, ex, t);
    group_sync<TPF>(bar);
    fwd_p3(br, bi, ex, tw, t);
  }
  __device__ static __forceinline__ void inv(float (&br)[16], float (&bi)[16], float (&ar)[16],
                                             float (&ai)[16], float2* ex, const float2* tw,
                                             const Regs& rg, int t, int bar) {
    inv_p1(br, bi, ex, tw, t);
    group_sync<TPF>(bar);
    inv_p2(br, bi, ex, t);
    group_sync<TPF>(bar);
    inv_p3(ar, ai, ex, rg, t);
  }
  // (pr,pi)[i] <- Z[N - (t + 256 i)]
  __device__ static __forceinline__ void mirror(const float (&br)[16], const float (&bi)[16],
                                                float (&pr)[16], float (&pi)[16], float2* ex, int t,
                                                int bar) {
    mirror_store(br, bi, ex, t);
    group_sync<TPF>(bar);
    mirror_load(pr, pi, ex, t);
  }
};

}  // namespace babe
