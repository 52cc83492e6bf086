// Power-of-two FFT of one constant-Q band held in registers: M = 256 * R3 points
// (R3 = 1, 2, 4, 8, 16 -> M = 256 ... 4096) by TPB = M / 16 threads, 16 points per
// thread, three passes 16 x 16 x R3 with two shared-memory exchanges:
//
//   n = 16 R3 n1 + R3 n2 + n3,   k = k1 + 16 k2 + 256 k3,   t = R3 n2 + n3
//   P1  thread t            FFT16 over n1 -> k1,  * W_M^{t k1}          -> ex[k1][t]
//   P2  thread (k1 = t & 15, n3 = t >> 4)   FFT16 over n2 -> k2, in place in ex
//   P3  thread t, pairs (k1 = t & 15, k2 = (t >> 4) + R3 j), j < 16 / R3:
//       * W_{16 R3}^{n3 k2},  FFT_R3 over n3 -> k3      =>  X[t + TPB j + 256 k3]
//
// ex rows hold TPB + 1 float2, which makes all three access patterns bank-conflict free
// (k1 runs fastest over the lanes in P2 and P3).  Input slot of register n1: TPB n1 + t;
// output slot of register j R3 + k3: t + TPB j + 256 k3 -- both are consecutive across
// the lanes, so the band's global loads and stores coalesce without staging.
// Every thread of the CTA must call fwd() (it contains CTA barriers).
#pragma once
#include "regfft.cuh"
#include "regfft_packed.cuh"

namespace babe {

template <int R3>
struct BandCore {
  static constexpr int M = 256 * R3, TPB = 16 * R3, ROW = TPB + 1, EX = 16 * ROW, NP = 16 / R3;
  struct Regs { float wr[16], wi[16]; };                      // W_M^{t k1}
  BABE_HD static void init_regs(Regs& r, const float2* roots_m, int t) {
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
      const float2 w = roots_m[t * k1];
      r.wr[k1] = w.x; r.wi[k1] = w.y;
    }
  }
  // tw (shared): W_{16 R3}^m = roots_m[16 m], m < TPB
  __device__ static __forceinline__ void load_twiddles(float2* tw, const float2* roots_m) {
    for (int i = threadIdx.x; i < TPB; i += blockDim.x) tw[i] = roots_m[16 * i];
  }
  // The three passes as per-thread functions (also callable on the host, where a test runs the
  // "threads" of one band one after the other: tests/host/bandfft_host_check.cu).
  BABE_HD static void pass1(float (&re)[16], float (&im)[16], float2* ex, const Regs& rg, int t) {
    fft16_split(re, im);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1)
      ex[k1 * ROW + t] = make_float2(re[k1] * rg.wr[k1] - im[k1] * rg.wi[k1],
                                     re[k1] * rg.wi[k1] + im[k1] * rg.wr[k1]);
  }
  // reads its column of ex, transforms it; for R3 > 1 the result goes back in place (call pass2_store
  // after ALL threads have run pass2_load when emulating on the host: the columns are disjoint, so the
  // device code stores immediately)
  BABE_HD static void pass2(float (&re)[16], float (&im)[16], float2* ex, int t) {
    float2* col = ex + (t & 15) * ROW + (t >> 4);            // + R3 n2
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) { const float2 v = col[R3 * n2]; re[n2] = v.x; im[n2] = v.y; }
    fft16_split(re, im);
    if (R3 == 1) return;                                      // X[t + 16 k2]
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) col[R3 * k2] = make_float2(re[k2], im[k2]);
  }
  BABE_HD static void pass3(float (&re)[16], float (&im)[16], const float2* ex, const float2* tw, int t) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const int k2 = (t >> 4) + R3 * j;
      const float2* row = ex + (t & 15) * ROW + R3 * k2;      // + n3
      float ar[R3], ai[R3];
#pragma unroll
      for (int n3 = 0; n3 < R3; ++n3) {
        float2 v = row[n3];
        if (n3 > 0) {
          const float2 w = tw[n3 * k2];
          v = make_float2(v.x * w.x - v.y * w.y, v.x * w.y + v.y * w.x);
        }
        ar[n3] = v.x; ai[n3] = v.y;
      }
      small_fft(ar, ai);
#pragma unroll
      for (int k3 = 0; k3 < R3; ++k3) { re[j * R3 + k3] = ar[k3]; im[j * R3 + k3] = ai[k3]; }
    }
  }
  __device__ static __forceinline__ void fwd(float (&re)[16], float (&im)[16], float2* ex,
                                             const float2* tw, const Regs& rg, int t) {
    pass1(re, im, ex, rg, t);
    __syncthreads();
    pass2(re, im, ex, t);
    if (R3 == 1) return;
    __syncthreads();
    pass3(re, im, ex, tw, t);
  }
  // slot of output register r
  BABE_HD static int out_slot(int r, int t) { return t + TPB * (r / R3) + 256 * (r % R3); }

 private:
  BABE_HD static void small_fft(float (&r)[1], float (&i)[1]) {}
  BABE_HD static void small_fft(float (&r)[2], float (&i)[2]) { fft2(r, i); }
  BABE_HD static void small_fft(float (&r)[4], float (&i)[4]) { fft4(r, i); }
  BABE_HD static void small_fft(float (&r)[8], float (&i)[8]) { fft8(r, i); }
  BABE_HD static void small_fft(float (&r)[16], float (&i)[16]) { fft16_split(r, i); }
};

}  // namespace babe
