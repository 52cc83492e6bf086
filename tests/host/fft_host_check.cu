// Host-side check of the __host__ __device__ FFT building blocks (regfft.cuh, regfft_packed.cuh,
// smemfft.cuh) against a naive double-precision DFT.  Compiled with nvcc and run on the CPU by
// tests/test_fft_host_cpu.py (no GPU needed); prints "name n max_rel_err" lines.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "regfft.cuh"
#include "regfft_packed.cuh"
#include "smemfft.cuh"

using namespace babe;

static double lcg_state = 12345.0;
static float rnd() {
  lcg_state = fmod(lcg_state * 1103515245.0 + 12345.0, 2147483648.0);
  return (float)(lcg_state / 2147483648.0 - 0.5);
}

static double check(const std::vector<float2>& in, const std::vector<float2>& out) {
  const int n = (int)in.size();
  double num = 0.0, den = 0.0;
  for (int k = 0; k < n; ++k) {
    double sr = 0.0, si = 0.0;
    for (int t = 0; t < n; ++t) {
      const double a = -2.0 * M_PI * (double)((long long)t * k % n) / n;
      sr += in[t].x * cos(a) - in[t].y * sin(a);
      si += in[t].x * sin(a) + in[t].y * cos(a);
    }
    num += (out[k].x - sr) * (out[k].x - sr) + (out[k].y - si) * (out[k].y - si);
    den += sr * sr + si * si;
  }
  return sqrt(num / den);
}

template <int R, class F>
static void reg_case(const char* name, F f) {
  std::vector<float2> in(R), out(R);
  float re[R], im[R];
  for (int i = 0; i < R; ++i) { in[i] = make_float2(rnd(), rnd()); re[i] = in[i].x; im[i] = in[i].y; }
  f(re, im);
  for (int i = 0; i < R; ++i) out[i] = make_float2(re[i], im[i]);
  printf("%s %d %.3e\n", name, R, check(in, out));
}

static void stockham_case(int n, std::vector<int> radices, int nseq) {
  FftFactors f{};
  f.n = n;
  f.nf = (int)radices.size();
  for (int s = 0; s < f.nf; ++s) f.radix[s] = radices[s];
  fill_fastdiv(f);
  const int S = padded_len(n);
  std::vector<float2> a((size_t)nseq * S), b((size_t)nseq * S), wn(n);
  for (int m = 0; m < n; ++m) wn[m] = make_float2((float)cos(-2.0 * M_PI * m / n), (float)sin(-2.0 * M_PI * m / n));
  std::vector<std::vector<float2>> in(nseq, std::vector<float2>(n));
  for (int q = 0; q < nseq; ++q)
    for (int i = 0; i < n; ++i) { in[q][i] = make_float2(rnd(), rnd()); a[(size_t)q * S + pad16(i)] = in[q][i]; }
  const float2* res = smem_fft(a.data(), b.data(), f, S, nseq, wn.data(), 0, 1);
  double worst = 0.0;
  for (int q = 0; q < nseq; ++q) {
    std::vector<float2> out(n);
    for (int i = 0; i < n; ++i) out[i] = res[(size_t)q * S + pad16(i)];
    worst = fmax(worst, check(in[q], out));
  }
  printf("stockham %d %.3e\n", n, worst);
}

int main() {
  reg_case<2>("fft2", [](float (&r)[2], float (&i)[2]) { fft2(r, i); });
  reg_case<4>("fft4", [](float (&r)[4], float (&i)[4]) { fft4(r, i); });
  reg_case<8>("fft8", [](float (&r)[8], float (&i)[8]) { fft8(r, i); });
  reg_case<16>("fft_reg", [](float (&r)[16], float (&i)[16]) { fft_reg<16>(r, i); });
  reg_case<32>("fft_reg", [](float (&r)[32], float (&i)[32]) { fft_reg<32>(r, i); });
  reg_case<64>("fft_reg", [](float (&r)[64], float (&i)[64]) { fft_reg<64>(r, i); });
  reg_case<16>("fft16_split", [](float (&r)[16], float (&i)[16]) { fft16_split(r, i); });
  reg_case<3>("dft_odd_sym", [](float (&r)[3], float (&i)[3]) { dft_odd_sym<3>(r, i); });
  reg_case<5>("dft_odd_sym", [](float (&r)[5], float (&i)[5]) { dft_odd_sym<5>(r, i); });
  reg_case<7>("dft_odd_sym", [](float (&r)[7], float (&i)[7]) { dft_odd_sym<7>(r, i); });
  reg_case<11>("dft_odd_sym", [](float (&r)[11], float (&i)[11]) { dft_odd_sym<11>(r, i); });
  reg_case<13>("dft_odd_sym", [](float (&r)[13], float (&i)[13]) { dft_odd_sym<13>(r, i); });
  reg_case<17>("dft_odd_sym", [](float (&r)[17], float (&i)[17]) { dft_odd_sym<17>(r, i); });
  reg_case<19>("dft_odd_sym", [](float (&r)[19], float (&i)[19]) { dft_odd_sym<19>(r, i); });
  reg_case<23>("dft_odd_sym", [](float (&r)[23], float (&i)[23]) { dft_odd_sym<23>(r, i); });
  // the pass lengths of the shipped segment lengths (184184 -> 308 x 299, 132300 -> 270 x 245,
  // 485100 -> 495 x 490, 368368 -> 616 x 299) and the small power-of-two octaves
  stockham_case(308, {4, 7, 11}, 2);
  stockham_case(299, {13, 23}, 2);
  stockham_case(270, {2, 3, 3, 3, 5}, 1);
  stockham_case(245, {5, 7, 7}, 1);
  stockham_case(495, {3, 3, 5, 11}, 1);
  stockham_case(490, {2, 5, 7, 7}, 1);
  stockham_case(616, {8, 7, 11}, 1);
  stockham_case(323, {17, 19}, 1);
  stockham_case(32, {16, 2}, 3);
  stockham_case(64, {16, 4}, 3);
  stockham_case(128, {16, 8}, 2);
  stockham_case(2048, {16, 16, 8}, 1);
  return 0;
}
