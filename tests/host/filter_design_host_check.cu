// Host-side run of the device filter-design logic (csrc/filter_design.cuh: breakpoint search,
// anchor chain, bin ownership, fp32 operation order of design_filter,
// utils/blind_bwe_utils.py:82-119).  Usage: prog in.bin out.bin
//   in : int32 F, int32 K, float f[F], float fc[K], float A[K] [, float gH[F]]
//   out: float H[F], int32 bad [, float gfc[K], float gA[K]  -- the VJP of the design for cotangent gH,
//        accumulated like k_design_filter_vjp and finished by finish_param_grads]
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "filter_design.cuh"

using namespace babe;

int main(int argc, char** argv) {
  if (argc != 3) return 2;
  FILE* fi = fopen(argv[1], "rb");
  if (!fi) return 3;
  int F = 0, K = 0;
  if (fread(&F, 4, 1, fi) != 1 || fread(&K, 4, 1, fi) != 1 || K > BABE_MAX_BREAKPOINTS) return 4;
  std::vector<float> f(F), fc(K), A(K), H(F);
  if (fread(f.data(), 4, F, fi) != (size_t)F || fread(fc.data(), 4, K, fi) != (size_t)K ||
      fread(A.data(), 4, K, fi) != (size_t)K) return 5;
  std::vector<float> gH(F);
  const bool vjp = fread(gH.data(), 4, F, fi) == (size_t)F;
  fclose(fi);
  FilterSegs segs;
  build_segments(segs, fc.data(), A.data(), K, f.data(), F);
  for (int k = 0; k < F; ++k) H[k] = bin_gain(segs, k, f[k]);
  FILE* fo = fopen(argv[2], "wb");
  fwrite(H.data(), 4, F, fo);
  fwrite(&segs.bad, 4, 1, fo);
  if (vjp) {
    double sv[BABE_MAX_BREAKPOINTS] = {0}, lv[BABE_MAX_BREAKPOINTS] = {0}, gfc[BABE_MAX_BREAKPOINTS], gA[BABE_MAX_BREAKPOINTS];
    for (int k = 0; k < F; ++k) {
      const int o = bin_owner(segs, k);
      if (o < 0) continue;
      const double u = (double)gH[k] * (double)H[k];
      sv[o] += u;
      lv[o] += u * (double)log2f(rn_div(f[k], segs.fc[o]));
    }
    finish_param_grads(segs, f.data(), F, sv, lv, gfc, gA);
    std::vector<float> o32(2 * K);
    for (int i = 0; i < K; ++i) { o32[i] = (float)gfc[i]; o32[K + i] = (float)gA[i]; }
    fwrite(o32.data(), 4, 2 * K, fo);
  }
  fclose(fo);
  return 0;
}
