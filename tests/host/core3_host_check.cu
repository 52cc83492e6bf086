// Host-side emulation of Core3 (csrc/stft_cores.cuh), the 4096-point transform core of
// k_apply_filter / k_stft_stats / k_fir_filter: the pass functions are the very ones the kernels call;
// the 256 "threads" of a frame-pair group run one after the other.  Prints
//   fwd  <rel err vs naive DFT>,  roundtrip <rel err of inv(fwd(z)) / N vs z>,  mirror <max abs err>
#include <math.h>
#include <stdio.h>

#include <vector>

#include "stft_cores.cuh"

using namespace babe;

static double lcg_state = 4242.0;
static float rnd() {
  lcg_state = fmod(lcg_state * 1103515245.0 + 12345.0, 2147483648.0);
  return (float)(lcg_state / 2147483648.0 - 0.5);
}

int main() {
  constexpr int N = Core3::N, T = Core3::TPF;
  std::vector<float2> z(N), Z(N), roots(N), tw(256), ex(Core3::EX_ELEMS), back(N);
  for (int m = 0; m < N; ++m) {
    z[m] = make_float2(rnd(), rnd());
    roots[m] = make_float2((float)cos(-2.0 * M_PI * m / N), (float)sin(-2.0 * M_PI * m / N));
  }
  for (int m = 0; m < 256; ++m) tw[m] = roots[16 * m];
  std::vector<Core3::Regs> rg(T);
  for (int t = 0; t < T; ++t) Core3::init_regs(rg[t], roots.data(), t);
  std::vector<float> br(16 * T), bi(16 * T);
  // ---- forward: thread t holds z[256 i + t] -> Z[t + 256 i]
  for (int t = 0; t < T; ++t) {
    float ar[16], ai[16];
    for (int i = 0; i < 16; ++i) { ar[i] = z[256 * i + t].x; ai[i] = z[256 * i + t].y; }
    Core3::fwd_p1(ar, ai, ex.data(), rg[t], t);
  }
  for (int t = 0; t < T; ++t) { float r[16], i[16]; Core3::fwd_p2(r, i, ex.data(), t); }
  for (int t = 0; t < T; ++t) {
    float r[16], i[16];
    Core3::fwd_p3(r, i, ex.data(), tw.data(), t);
    for (int q = 0; q < 16; ++q) { br[16 * t + q] = r[q]; bi[16 * t + q] = i[q]; Z[t + 256 * q] = make_float2(r[q], i[q]); }
  }
  double num = 0.0, den = 0.0;
  for (int k = 0; k < N; k += 7) {                               // every 7th bin: 586 naive sums
    double sr = 0.0, si = 0.0;
    for (int n = 0; n < N; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)n * k % N) / N;
      sr += z[n].x * cos(a) - z[n].y * sin(a);
      si += z[n].x * sin(a) + z[n].y * cos(a);
    }
    num += (Z[k].x - sr) * (Z[k].x - sr) + (Z[k].y - si) * (Z[k].y - si);
    den += sr * sr + si * si;
  }
  printf("fwd %.3e\n", sqrt(num / den));
  // ---- mirror: (pr, pi)[i] = Z[N - (t + 256 i)]
  for (int t = 0; t < T; ++t) {
    float r[16], i[16];
    for (int q = 0; q < 16; ++q) { r[q] = br[16 * t + q]; i[q] = bi[16 * t + q]; }
    Core3::mirror_store(r, i, ex.data(), t);
  }
  double merr = 0.0;
  for (int t = 0; t < T; ++t) {
    float pr[16], pi[16];
    Core3::mirror_load(pr, pi, ex.data(), t);
    for (int q = 0; q < 16; ++q) {
      const float2 want = Z[(N - (t + 256 * q)) % N];
      merr = fmax(merr, fmax(fabs(pr[q] - want.x), fabs(pi[q] - want.y)));
    }
  }
  printf("mirror %.3e\n", merr);
  // ---- inverse: Z[t + 256 i] -> N z[256 i + t]
  for (int t = 0; t < T; ++t) {
    float r[16], i[16];
    for (int q = 0; q < 16; ++q) { r[q] = br[16 * t + q]; i[q] = bi[16 * t + q]; }
    Core3::inv_p1(r, i, ex.data(), tw.data(), t);
  }
  for (int t = 0; t < T; ++t) { float r[16], i[16]; Core3::inv_p2(r, i, ex.data(), t); }
  num = den = 0.0;
  for (int t = 0; t < T; ++t) {
    float ar[16], ai[16];
    Core3::inv_p3(ar, ai, ex.data(), rg[t], t);
    for (int i = 0; i < 16; ++i) {
      const float2 want = z[256 * i + t];
      const double dx = ar[i] / (double)N - want.x, dy = ai[i] / (double)N - want.y;
      num += dx * dx + dy * dy;
      den += (double)want.x * want.x + (double)want.y * want.y;
    }
  }
  printf("roundtrip %.3e\n", sqrt(num / den));
  return 0;
}
