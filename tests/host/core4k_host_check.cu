// Host-side emulation of Core4k (csrc/core4k.cuh), the 4096-point transform core of the fused
// STFT-domain kernels (csrc/stft_fused.cu): the pass functions are the very ones the kernels call; the
// 256 "threads" of a frame-pair group run one after the other, with the load / store halves of the
// in-place passes separated exactly where the device code has its __syncwarp().  Prints
//   fwd <rel err vs naive DFT>   roundtrip <rel err of inv(fwd(z)) / N>   perm <max abs err of the
//   permuted-table indexing>   fft16 <rel err of fft16v fwd and inv vs naive>
#include <math.h>
#include <stdio.h>

#include <vector>

#include "core4k.cuh"

using namespace babe;

static double lcg_state = 777.0;
static float rnd() {
  lcg_state = fmod(lcg_state * 1103515245.0 + 12345.0, 2147483648.0);
  return (float)(lcg_state / 2147483648.0 - 0.5);
}

template <class TW>
static void run(const std::vector<float2>& roots, const std::vector<float2>& tw4tab, bool smem_tw) {
  constexpr int N = Core4k::N, T = Core4k::TPF;
  std::vector<float2> z(N), Z(N), tw3(256), ex(Core4k::EX);
  for (int m = 0; m < N; ++m) z[m] = make_float2(rnd(), rnd());
  for (int n3 = 0; n3 < 16; ++n3)
    for (int k2 = 0; k2 < 16; ++k2) tw3[16 * n3 + k2] = roots[(16 * n3 * k2) % N];
  std::vector<TW> tw(T);
  for (int t = 0; t < T; ++t) tw[t].init(smem_tw ? tw4tab.data() : roots.data(), t);
  std::vector<float2> regs(16 * T), tmp(16 * T);
  // forward
  for (int t = 0; t < T; ++t) {
    float2 v[16];
    for (int i = 0; i < 16; ++i) v[i] = z[256 * i + t];
    Core4k::fwd_p1(v, ex.data(), tw[t], t);
  }
  for (int t = 0; t < T; ++t) { float2 v[16]; Core4k::fwd_p2_load(v, ex.data(), t); for (int i = 0; i < 16; ++i) tmp[16 * t + i] = v[i]; }
  for (int t = 0; t < T; ++t) { float2 v[16]; for (int i = 0; i < 16; ++i) v[i] = tmp[16 * t + i]; Core4k::fwd_p2_store(v, ex.data(), t); }
  for (int t = 0; t < T; ++t) {
    float2 v[16];
    Core4k::fwd_p3(v, ex.data(), tw3.data(), t);
    for (int q = 0; q < 16; ++q) { regs[16 * t + q] = v[q]; Z[Core4k::bin_of(t, q)] = v[q]; }
  }
  double num = 0.0, den = 0.0;
  for (int k = 0; k < N; k += 7) {
    double sr = 0.0, si = 0.0;
    for (int n = 0; n < N; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)n * k % N) / N;
      sr += z[n].x * cos(a) - z[n].y * sin(a);
      si += z[n].x * sin(a) + z[n].y * cos(a);
    }
    num += (Z[k].x - sr) * (Z[k].x - sr) + (Z[k].y - si) * (Z[k].y - si);
    den += sr * sr + si * si;
  }
  printf("fwd_%s %.3e\n", smem_tw ? "smem" : "regs", sqrt(num / den));
  // inverse
  for (int t = 0; t < T; ++t) {
    float2 v[16];
    for (int q = 0; q < 16; ++q) v[q] = regs[16 * t + q];
    Core4k::inv_q1(v, ex.data(), tw3.data(), t);
  }
  for (int t = 0; t < T; ++t) { float2 v[16]; Core4k::inv_q2_load(v, ex.data(), t); for (int i = 0; i < 16; ++i) tmp[16 * t + i] = v[i]; }
  for (int t = 0; t < T; ++t) { float2 v[16]; for (int i = 0; i < 16; ++i) v[i] = tmp[16 * t + i]; Core4k::inv_q2_store(v, ex.data(), t); }
  num = den = 0.0;
  for (int t = 0; t < T; ++t) {
    float2 v[16];
    Core4k::inv_q3(v, ex.data(), tw[t], t);
    for (int i = 0; i < 16; ++i) {
      const float2 want = z[256 * i + t];
      const double dx = v[i].x / (double)N - want.x, dy = v[i].y / (double)N - want.y;
      num += dx * dx + dy * dy;
      den += (double)want.x * want.x + (double)want.y * want.y;
    }
  }
  printf("roundtrip_%s %.3e\n", smem_tw ? "smem" : "regs", sqrt(num / den));
}

int main() {
  constexpr int N = Core4k::N, T = Core4k::TPF;
  std::vector<float2> roots(N), tw4(15 * 256);
  for (int m = 0; m < N; ++m) roots[m] = make_float2((float)cos(-2.0 * M_PI * m / N), (float)sin(-2.0 * M_PI * m / N));
  for (int k1 = 1; k1 < 16; ++k1)
    for (int t = 0; t < 256; ++t) tw4[(k1 - 1) * 256 + t] = roots[t * k1];
  // fft16v forward / inverse vs naive
  {
    float2 v[16], w[16];
    double num = 0, den = 0;
    for (int i = 0; i < 16; ++i) v[i] = w[i] = make_float2(rnd(), rnd());
    float2 a[16], b[16];
    for (int i = 0; i < 16; ++i) { a[i] = v[i]; b[i] = v[i]; }
    fft16v<false>(a);
    fft16v<true>(b);
    for (int k = 0; k < 16; ++k) {
      double fr = 0, fi = 0, ir = 0, ii = 0;
      for (int n = 0; n < 16; ++n) {
        const double c = cos(2 * M_PI * n * k / 16), s = sin(2 * M_PI * n * k / 16);
        fr += v[n].x * c + v[n].y * s; fi += v[n].y * c - v[n].x * s;
        ir += v[n].x * c - v[n].y * s; ii += v[n].y * c + v[n].x * s;
      }
      num += (a[k].x - fr) * (a[k].x - fr) + (a[k].y - fi) * (a[k].y - fi) + (b[k].x - ir) * (b[k].x - ir) + (b[k].y - ii) * (b[k].y - ii);
      den += fr * fr + fi * fi + ir * ir + ii * ii;
    }
    printf("fft16 %.3e\n", sqrt(num / den));
    // scaled first stage == transform of the pre-scaled values
    float sc[16];
    float2 c[16], d[16];
    double n2 = 0, d2 = 0;
    for (int i = 0; i < 16; ++i) { sc[i] = 0.5f + rnd(); c[i] = v[i]; d[i] = make_float2(v[i].x * sc[i], v[i].y * sc[i]); }
    fft16v_scaled<true>(c, sc);
    fft16v<true>(d);
    for (int i = 0; i < 16; ++i) { n2 += (c[i].x - d[i].x) * (c[i].x - d[i].x) + (c[i].y - d[i].y) * (c[i].y - d[i].y); d2 += d[i].x * d[i].x + d[i].y * d[i].y; }
    printf("fft16_scaled %.3e\n", sqrt(n2 / d2));
  }
  run<Core4k::TwRegs>(roots, tw4, false);
  run<Core4k::TwSmem>(roots, tw4, true);
  // permuted real-symmetric table: entry of bin k and of its mirror N - k as the threads address them
  std::vector<float> hs(2049), hp(2049);
  for (int k = 0; k <= 2048; ++k) hs[k] = (float)(k + 1);
  for (int k = 0; k <= 2048; ++k) hp[Core4k::perm_of_bin(k)] = hs[k];
  double perr = 0.0;
  for (int t = 0; t < T; ++t) {
    const int pm = Core4k::mirror_base(t);
    for (int k3 = 0; k3 < 16; ++k3) {
      const int k = Core4k::bin_of(t, k3);
      const float want = hs[k <= 2048 ? k : N - k];
      const float got = (k3 < 8) ? hp[256 * k3 + t] : hp[256 * (15 - k3) + pm];
      perr = fmax(perr, fabs(got - want));
    }
  }
  printf("perm %.3e\n", perr);
  return 0;
}
