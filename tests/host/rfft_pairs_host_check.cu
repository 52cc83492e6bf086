// Host-side check of the r2c / c2r bin-pair processing (csrc/rfft_pairs.cuh) that turns the complex
// FFT of Nc = Ls/2 points into the real FFT of Ls points and back (k_rfft_post, k_irfft_pre,
// k_spectral_mid, k_cqt_gather_pre).  Prints "post <rel err vs naive real DFT>" and
// "pre <rel err of the recovered conj(Z)/Nc>".
#include <math.h>
#include <stdio.h>

#include <vector>

#include "rfft_pairs.cuh"

using namespace babe;

int main() {
  const int Nc = 2 * 3 * 5 * 7, Ls = 2 * Nc;
  std::vector<double> x(Ls);
  double st = 99.0;
  for (int n = 0; n < Ls; ++n) { st = fmod(st * 1103515245.0 + 12345.0, 2147483648.0); x[n] = st / 2147483648.0 - 0.5; }
  std::vector<float2> Z(Nc), X(Nc + 1), W(Nc + 1);
  for (int k = 0; k < Nc; ++k) {                                 // Z = FFT_Nc(x_even + i x_odd), naive
    double sr = 0.0, si = 0.0;
    for (int n = 0; n < Nc; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)n * k % Nc) / Nc, c = cos(a), s = sin(a);
      sr += x[2 * n] * c - x[2 * n + 1] * s;
      si += x[2 * n] * s + x[2 * n + 1] * c;
    }
    Z[k] = make_float2((float)sr, (float)si);
  }
  for (int k = 0; k <= Nc; ++k) W[k] = make_float2((float)cos(-2.0 * M_PI * k / Ls), (float)sin(-2.0 * M_PI * k / Ls));
  X[0] = make_float2(Z[0].x + Z[0].y, 0.f);                      // as in k_rfft_post
  X[Nc] = make_float2(Z[0].x - Z[0].y, 0.f);
  for (int k = 1; k <= Nc / 2; ++k) {
    float2 a, b;
    post_pair(Z[k], Z[Nc - k], W[k], a, b);
    X[k] = a;
    if (Nc - k != k) X[Nc - k] = b;
  }
  double num = 0.0, den = 0.0;
  for (int k = 0; k <= Nc; ++k) {
    double sr = 0.0, si = 0.0;
    for (int n = 0; n < Ls; ++n) {
      const double a = -2.0 * M_PI * (double)((long long)n * k % Ls) / Ls;
      sr += x[n] * cos(a); si += x[n] * sin(a);
    }
    num += (X[k].x - sr) * (X[k].x - sr) + (X[k].y - si) * (X[k].y - si);
    den += sr * sr + si * si;
  }
  printf("post %.3e\n", sqrt(num / den));
  // c2r pre-processing: X -> conj(Z[k]) / Nc
  num = den = 0.0;
  for (int k = 0; k <= Nc / 2; ++k) {
    const int kp = Nc - k;
    float2 a = X[k], b = X[kp], zk, zkp;
    if (k == 0) { a.y = 0.f; b.y = 0.f; }
    pre_pair(a, b, W[k], 1.0f / (float)Nc, zk, zkp);
    const int ks[2] = {k, kp % Nc};
    const float2 got[2] = {zk, zkp};
    for (int q = 0; q < ((k == 0 || kp == k) ? 1 : 2); ++q) {
      const double wr = Z[ks[q]].x / (double)Nc, wi = -Z[ks[q]].y / (double)Nc;
      num += (got[q].x - wr) * (got[q].x - wr) + (got[q].y - wi) * (got[q].y - wi);
      den += wr * wr + wi * wi;
    }
  }
  printf("pre %.3e\n", sqrt(num / den));
  return 0;
}
