// Host-side emulation of BandCoreV<R3, INV> and BandCoreS<R2, INV> (csrc/bandfft_v.cuh): the per-thread passes the
// CQT band kernels call, run for the "threads" of one band one after the other, forward and inverse, with and
// without the folded window multiply.  Prints "<core> <M> <inv> <scaled> <rel err vs naive double DFT>".
#include <math.h>
#include <stdio.h>

#include <vector>

#include "bandfft_v.cuh"

using namespace babe;

static double lcg_state = 4242.0;
static float rnd() {
  lcg_state = fmod(lcg_state * 1103515245.0 + 12345.0, 2147483648.0);
  return (float)(lcg_state / 2147483648.0 - 0.5);
}

template <class C, bool INV, bool SCALED, bool THREE>
static void run(const char* name) {
  constexpr int M = C::M, TPB = C::TPB;
  std::vector<float2> x(M), out(M), roots(M), tw(C::NTW), ex(C::EX);
  std::vector<float> s(M);
  for (int m = 0; m < M; ++m) {
    x[m] = make_float2(rnd(), rnd());
    s[m] = 0.5f + rnd();
    roots[m] = make_float2((float)cos(-2.0 * M_PI * m / M), (float)sin(-2.0 * M_PI * m / M));
  }
  for (int m = 0; m < C::NTW; ++m) tw[m] = C::twiddle(roots.data(), m);
  std::vector<float2> keep(16 * TPB);
  for (int t = 0; t < TPB; ++t) {
    float2 z[16];
    float sc[16];
    for (int n1 = 0; n1 < 16; ++n1) { z[n1] = x[C::in_slot(n1, t)]; sc[n1] = s[C::in_slot(n1, t)]; }
    typename C::Regs rg;
    C::init_regs(rg, roots.data(), t);
    C::template pass1<SCALED>(z, sc, ex.data(), rg, t);
  }
  for (int t = 0; t < TPB; ++t) {
    float2 z[16];
    C::pass2(z, ex.data(), t);
    for (int q = 0; q < 16; ++q) keep[16 * t + q] = z[q];
  }
  for (int t = 0; t < TPB; ++t) {
    float2 z[16];
    for (int q = 0; q < 16; ++q) z[q] = keep[16 * t + q];
    if constexpr (THREE) { if (C::NP < 16) C::pass3(z, ex.data(), tw.data(), t); }
    for (int q = 0; q < 16; ++q) out[C::out_slot(q, t)] = z[q];
  }
  double num = 0.0, den = 0.0;
  const double sign = INV ? 2.0 : -2.0;
  for (int k = 0; k < M; ++k) {
    double sr = 0.0, si = 0.0;
    for (int n = 0; n < M; ++n) {
      const double a = sign * M_PI * (double)((long long)n * k % M) / M;
      const double xr = x[n].x * (SCALED ? s[n] : 1.0), xi = x[n].y * (SCALED ? s[n] : 1.0);
      sr += xr * cos(a) - xi * sin(a);
      si += xr * sin(a) + xi * cos(a);
    }
    num += (out[k].x - sr) * (out[k].x - sr) + (out[k].y - si) * (out[k].y - si);
    den += sr * sr + si * si;
  }
  printf("%s %d %d %d %.3e\n", name, M, (int)INV, (int)SCALED, sqrt(num / den));
}

template <int R3> static void run_v() {
  run<BandCoreV<R3, false>, false, false, true>("V");
  run<BandCoreV<R3, true>, true, true, true>("V");
}
template <int R2> static void run_s() {
  run<BandCoreS<R2, false>, false, false, false>("S");
  run<BandCoreS<R2, true>, true, true, false>("S");
}

int main() {
  run_s<2>(); run_s<4>(); run_s<8>();
  run_v<1>(); run_v<2>(); run_v<4>(); run_v<8>(); run_v<16>();
  return 0;
}
