// 4096-point complex FFT / inverse FFT of one frame pair held in the registers of a 256-thread group
// (16 float2 per thread), three radix-16 passes, TWO exchanges through shared memory per direction of
// which only ONE needs a group-wide barrier:
//
//   n = 256 n1 + 16 n2 + n3,   k = k1 + 16 k2 + 256 k3
//   P1  thread t = 16 n2 + n3      FFT16 n1 -> k1,  * W_4096^{t k1}      -> ex[k1][t]         (row pitch 272)
//       ---- group barrier ----
//   P2  thread t = 16 k1 + n3      FFT16 n2 -> k2, back IN PLACE into its own row as  ex[k1][17 k2 + n3]
//       ---- __syncwarp: the 16 threads of a row are one half-warp ----
//   P3  thread t = 16 k1 + k2      * W_256^{n3 k2},  FFT16 n3 -> k3      => Z[k1 + 16 k2 + 256 k3] in register k3
//
// The inverse runs the passes backwards (Q1, Q2, Q3) with conjugated twiddles and ends with thread t holding
// N * z[256 n1 + t] in register n1.  All shared-memory accesses are 64-bit and conflict-free: consecutive lanes
// touch consecutive float2 (P1 / P2 loads / Q2 / Q3) or float2 at stride 17 (P2 stores / P3 / Q1: 34 words,
// i.e. 16 distinct even banks per half-warp).  Round 1's Core3 (stft_cores.cuh) used a [16][257] buffer with
// four group barriers per transform pair; here it is two.
//
// The per-thread pass functions are __host__ __device__: tests/host/core4k_host_check.cu emulates the 256
// threads on the CPU.  On the device, `load` and `store` halves of P2 / Q2 must be separated by __syncwarp().
#pragma once
#include "fft16v.cuh"

namespace babe {

struct Core4k {
  static constexpr int N = 4096, HOP = 2048, F = 2049, TPF = 256;
  static constexpr int PITCH = 272, EX = 16 * PITCH;          // float2 elements per group
  // twiddle tables (float2), both derived from roots[m] = exp(-2 pi i m / 4096):
  //   tw3[16 n3 + k2] = W_256^{n3 k2} = roots[16 n3 k2]                       (shared memory, 2 KB)
  //   W_4096^{t k1} = roots[t k1]: 16 per thread, in registers (TwRegs) or in a [k1][t] shared table (TwSmem)
  struct TwRegs {
    float2 w[16];
    BABE_HD void init(const float2* roots, int t) {
#pragma unroll
      for (int k1 = 0; k1 < 16; ++k1) w[k1] = roots[t * k1];
    }
    BABE_HD float2 get(int k1) const { return w[k1]; }
  };
  struct TwSmem {                       // table rows k1 = 1..15 (row 0 is all ones)
    const float2* p;
    BABE_HD void init(const float2* table, int t) { p = table + t - 256; }
    BABE_HD float2 get(int k1) const { return p[k1 * 256]; }
  };

  // ---- forward -----------------------------------------------------------------------------------
  // SCALED: the transform of z[n1] * s[n1] (the analysis window rides on the first butterflies)
  template <class TW, bool SCALED = false>
  BABE_HD static void fwd_p1(float2 (&z)[16], float2* ex, const TW& tw, int t, const float (&s)[16]) {
    fft16v_impl<false, SCALED>(z, s);
    ex[t] = z[0];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) ex[k1 * PITCH + t] = c_mul(z[k1], tw.get(k1));
  }
  template <class TW>
  BABE_HD static void fwd_p1(float2 (&z)[16], float2* ex, const TW& tw, int t) {
    const float none[16] = {};
    fwd_p1<TW, false>(z, ex, tw, t, none);
  }
  BABE_HD static void fwd_p2_load(float2 (&v)[16], const float2* ex, int t) {
    const float2* row = ex + (t >> 4) * PITCH + (t & 15);
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) v[n2] = row[16 * n2];
  }
  BABE_HD static void fwd_p2_store(float2 (&v)[16], float2* ex, int t) {
    fft16v<false>(v);
    float2* row = ex + (t >> 4) * PITCH + (t & 15);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) row[17 * k2] = v[k2];
  }
  BABE_HD static void fwd_p3(float2 (&v)[16], const float2* ex, const float2* tw3, int t) {
    const float2* row = ex + (t >> 4) * PITCH + 17 * (t & 15);
    const float2* w = tw3 + (t & 15);
    v[0] = row[0];
#pragma unroll
    for (int n3 = 1; n3 < 16; ++n3) v[n3] = c_mul(row[n3], w[16 * n3]);
    fft16v<false>(v);
  }
  // ---- inverse (unnormalised) -----------------------------------------------------------------------
  // SCALED: the inverse transform of v[k3] * s[k3] (the filter gain rides on the first butterflies)
  template <bool SCALED = false>
  BABE_HD static void inv_q1(float2 (&v)[16], float2* ex, const float2* tw3, int t, const float (&s)[16]) {
    fft16v_impl<true, SCALED>(v, s);
    float2* row = ex + (t >> 4) * PITCH + 17 * (t & 15);
    const float2* w = tw3 + (t & 15);
    row[0] = v[0];
#pragma unroll
    for (int n3 = 1; n3 < 16; ++n3) row[n3] = c_mulc(v[n3], w[16 * n3]);
  }
  BABE_HD static void inv_q1(float2 (&v)[16], float2* ex, const float2* tw3, int t) {
    const float none[16] = {};
    inv_q1<false>(v, ex, tw3, t, none);
  }
  BABE_HD static void inv_q2_load(float2 (&v)[16], const float2* ex, int t) {
    const float2* row = ex + (t >> 4) * PITCH + (t & 15);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) v[k2] = row[17 * k2];
  }
  BABE_HD static void inv_q2_store(float2 (&v)[16], float2* ex, int t) {
    fft16v<true>(v);
    float2* row = ex + (t >> 4) * PITCH + (t & 15);
#pragma unroll
    for (int n2 = 0; n2 < 16; ++n2) row[16 * n2] = v[n2];
  }
  template <class TW>
  BABE_HD static void inv_q3(float2 (&z)[16], const float2* ex, const TW& tw, int t) {
    z[0] = ex[t];
#pragma unroll
    for (int k1 = 1; k1 < 16; ++k1) z[k1] = c_mulc(ex[k1 * PITCH + t], tw.get(k1));
    fft16v<true>(z);
  }
  // bin held in register k3 of thread t after fwd_p3
  BABE_HD static int bin_of(int t, int k3) { return (t >> 4) + 16 * (t & 15) + 256 * k3; }
  // position of bin k (< 2048 + 1) in a table permuted so that thread t reads entry 256 k3 + t for its
  // register k3 < 8 (and the Nyquist bin sits at 2048)
  BABE_HD static int perm_of_bin(int k) {
    if (k >= 2048) return 2048;
    const int k3 = k >> 8, k2 = (k >> 4) & 15, k1 = k & 15;
    return 256 * k3 + 16 * k1 + k2;
  }
  // For registers k3 >= 8 the bin N - k of a real-symmetric table is needed: it sits at 256 (15 - k3) + pm
  // with the per-thread constant pm below (256 for thread 0, whose mirror bins are the multiples of 256).
  BABE_HD static int mirror_base(int t) {
    if (t == 0) return 256;
    const int q = 256 - ((t >> 4) + 16 * (t & 15));     // 256 - (k1 + 16 k2) in [1, 255]
    return 16 * (q & 15) + (q >> 4);
  }
};

}  // namespace babe
