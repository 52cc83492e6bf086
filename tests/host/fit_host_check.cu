// Host-side emulation of the filter-fit iteration (k_fit_params, csrc/fit_ops.cu) built from the
// device functions of csrc/filter_design.cuh: segments, per-bin gain, chain rule, fp32 step,
// sequential clamps, stopping test -- summed serially instead of by 16 warps.
// Usage: prog in.bin out.bin
//   in : int32 F, K; babe_fit_config; double abc[3F]; float w[F]; float f[F]; float params[2K]
//   out: float params[2K]; int32 iterations
#include <math.h>
#include <stdio.h>

#include <vector>

#include "filter_design.cuh"

using namespace babe;

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) return 3;
  int F = 0, K = 0;
  babe_fit_config cfg;
  if (fread(&F, 4, 1, fi) != 1 || fread(&K, 4, 1, fi) != 1 || fread(&cfg, sizeof(cfg), 1, fi) != 1) return 4;
  std::vector<double> abc(3 * F);
  std::vector<float> w(F), f(F), p(2 * K);
  if (fread(abc.data(), 8, 3 * F, fi) != (size_t)(3 * F) || fread(w.data(), 4, F, fi) != (size_t)F ||
      fread(f.data(), 4, F, fi) != (size_t)F || fread(p.data(), 4, 2 * K, fi) != (size_t)(2 * K)) return 5;
  fclose(fi);
  std::vector<double> wa(F), wb(F);
  double c_total = 0.0;
  for (int k = 0; k < F; ++k) {
    const double w2 = (double)w[k] * (double)w[k];
    wa[k] = w2 * abc[k]; wb[k] = w2 * abc[F + k]; c_total += w2 * abc[2 * F + k];
  }
  std::vector<float> fc(p.begin(), p.begin() + K), A(p.begin() + K, p.end()), fc_prev(K), A_prev(K);
  int it = 0;
  for (int iter = 0; iter < cfg.max_iter; ++iter) {
    FilterSegs segs;
    build_segments(segs, fc.data(), A.data(), K, f.data(), F);
    double sv[BABE_MAX_BREAKPOINTS] = {0}, lv[BABE_MAX_BREAKPOINTS] = {0}, loss = 0.0;
    for (int k = 0; k < F; ++k) {
      const int o = bin_owner(segs, k);
      const double h = (double)bin_gain(segs, k, f[k]);
      loss += h * (h * wa[k] - 2.0 * wb[k]);
      if (o >= 0) {
        const double u = (h * wa[k] - wb[k]) * h;
        sv[o] += u;
        lv[o] += u * (double)log2f(rn_div(f[k], segs.fc[o]));
      }
    }
    double gfc[BABE_MAX_BREAKPOINTS], gA[BABE_MAX_BREAKPOINTS];
    finish_param_grads(segs, f.data(), F, sv, lv, gfc, gA);
    const double S = loss + c_total;
    const double inv_norm = 1.0 / sqrt(S > 0.0 ? S : 0.0);
    for (int j = 0; j < K; ++j) {                              // fp32 step like the reference (:569)
      fc[j] = rn_sub(fc[j], rn_mul(cfg.mu_fc, (float)(gfc[j] * inv_norm)));
      A[j] = rn_sub(A[j], rn_mul(cfg.mu_A, (float)(gA[j] * inv_norm)));
    }
    fit_project(fc.data(), A.data(), K, cfg);
    const bool stop = iter > 0 && fit_converged(fc.data(), A.data(), fc_prev.data(), A_prev.data(), K, cfg);
    fc_prev = fc; A_prev = A;
    it = iter + 1;
    if (stop) break;
  }
  FILE* fo = fopen(argv[2], "wb");
  fwrite(fc.data(), 4, K, fo);
  fwrite(A.data(), 4, K, fo);
  fwrite(&it, 4, 1, fo);
  fclose(fo);
  return 0;
}
