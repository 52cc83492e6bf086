// Host-side emulation of BandCore<R3>::fwd (csrc/bandfft.cuh): the three passes are the very
// functions the CUDA kernels call; here the "threads" of one band run one after the other, with the
// shared exchange buffer in host memory.  Prints "bandcore M max_rel_err" per size.
#include <math.h>
#include <stdio.h>

#include <vector>

#include "bandfft.cuh"

using namespace babe;

static double lcg_state = 777.0;
static float rnd() {
  lcg_state = fmod(lcg_state * 1103515245.0 + 12345.0, 2147483648.0);
  return (float)(lcg_state / 2147483648.0 - 0.5);
}

template <int R3>
static void run() {
  using C = BandCore<R3>;
  constexpr int M = C::M, TPB = C::TPB;
  std::vector<float2> x(M), out(M), roots(M), tw(TPB), ex(C::EX);
  for (int m = 0; m < M; ++m) {
    x[m] = make_float2(rnd(), rnd());
    roots[m] = make_float2((float)cos(-2.0 * M_PI * m / M), (float)sin(-2.0 * M_PI * m / M));
  }
  for (int m = 0; m < TPB; ++m) tw[m] = roots[16 * m];
  std::vector<float> re(16 * TPB), im(16 * TPB);
  for (int t = 0; t < TPB; ++t) {                               // pass 1, every thread
    float r[16], i[16];
    for (int n1 = 0; n1 < 16; ++n1) { r[n1] = x[TPB * n1 + t].x; i[n1] = x[TPB * n1 + t].y; }
    typename C::Regs rg;
    C::init_regs(rg, roots.data(), t);
    C::pass1(r, i, ex.data(), rg, t);
  }
  for (int t = 0; t < TPB; ++t) {                               // barrier; pass 2 (disjoint columns)
    float r[16], i[16];
    C::pass2(r, i, ex.data(), t);
    for (int q = 0; q < 16; ++q) { re[16 * t + q] = r[q]; im[16 * t + q] = i[q]; }
  }
  for (int t = 0; t < TPB; ++t) {                               // barrier; pass 3
    float r[16], i[16];
    for (int q = 0; q < 16; ++q) { r[q] = re[16 * t + q]; i[q] = im[16 * t + q]; }
    if (R3 > 1) C::pass3(r, i, ex.data(), tw.data(), t);
    for (int q = 0; q < 16; ++q) out[C::out_slot(q, t)] = make_float2(r[q], i[q]);
  }
  double num = 0.0, den = 0.0;
  for (int k = 0; k < M; ++k) {
    double sr = 0.0, si = 0.0;
    for (int n = 0; n < M; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)n * k % M) / M;
      sr += x[n].x * cos(a) - x[n].y * sin(a);
      si += x[n].x * sin(a) + x[n].y * cos(a);
    }
    num += (out[k].x - sr) * (out[k].x - sr) + (out[k].y - si) * (out[k].y - si);
    den += sr * sr + si * si;
  }
  printf("bandcore %d %.3e\n", M, sqrt(num / den));
}

int main() {
  run<1>();
  run<2>();
  run<4>();
  run<8>();
  run<16>();
  return 0;
}
