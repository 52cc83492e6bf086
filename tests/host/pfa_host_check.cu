// Host emulation of the prime-factor passes of csrc/cqt_pfa.cuh: every per-thread phase of k_pfa1_fwd, k_pfa2_fwd,
// k_pfa2_mid, k_pfa2_inv (plain and gather) and k_pfa1_inv is run for all threads of all CTAs of one row, barrier by
// barrier.  Writes x, H, scale, the gather inputs and the four results as raw float32 files into argv[2];
// tests/test_fft_host_cpu.py compares them with numpy.   usage: pfa_host_check <plan: 0..7> <dir>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "cqt_pfa.cuh"

using namespace babe;
using namespace babe::pfa;

static double lcg_state = 12345.0;
static float rnd() {
  lcg_state = fmod(lcg_state * 1103515245.0 + 12345.0, 2147483648.0);
  return (float)(lcg_state / 2147483648.0 - 0.5);
}
static void dump(const std::string& dir, const char* name, const void* p, size_t bytes) {
  FILE* f = fopen((dir + "/" + name).c_str(), "wb");
  if (!f) { perror(name); exit(2); }
  fwrite(p, 1, bytes, f);
  fclose(f);
}

template <class PL, int S>
static void run(const std::string& dir) {
  using P1 = Pass1<PL, S>;
  using P2 = Pass2<PL, S>;
  const int Nc = PL::NC, Ls = 2 * Nc;
  std::vector<float> x(Ls), Hf(Nc + 1), scale(Nc + 1);
  for (auto& v : x) v = rnd();
  for (auto& v : Hf) v = 0.5f + rnd();
  for (auto& v : scale) v = 1.0f + 0.5f * rnd();
  std::vector<float2> twls(TW_LO_ + (Ls >> 10) + 1);
  for (int m = 0; m < TW_LO_; ++m) twls[m] = make_float2((float)cos(-2.0 * M_PI * m / Ls), (float)sin(-2.0 * M_PI * m / Ls));
  for (int m = 0; m <= (Ls >> 10); ++m)
    twls[TW_LO_ + m] = make_float2((float)cos(-2.0 * M_PI * 1024.0 * m / Ls), (float)sin(-2.0 * M_PI * 1024.0 * m / Ls));
  std::vector<float2> Y((size_t)PL::N1 * PL::P2), Y2((size_t)PL::N1 * PL::P2), X(Nc + 1), y(Nc), xr(Nc), xg(Nc);
  std::vector<unsigned char> smem1(P1::SMEM), smem2(P2::SMEM);
  auto pass1_fwd = [&](const float2* in, float2* Yo) {
    for (int tile = 0; tile < P1::TILES; ++tile) {
      float2* A = reinterpret_cast<float2*>(smem1.data());
      unsigned short* T1 = P1::table(A);
      for (int t = 0; t < THREADS; ++t) P1::tables(T1, t);
      for (int t = 0; t < THREADS; ++t) P1::load_natural(in, A, T1, tile, t);
      // the stages of forward_to_rows, barrier by barrier (the last one writes the rows of Y itself)
      for (int t = 0; t < THREADS; ++t) P1::template stage_n<false, 0, P1::NST == 1 ? 2 : 0>(A, nullptr, Yo, tile, t);
      for (int t = 0; t < THREADS; ++t) P1::template stage_n<false, 1, P1::NST == 2 ? 2 : 0>(A, nullptr, Yo, tile, t);
      for (int t = 0; t < THREADS; ++t) P1::template stage_n<false, 2, 2>(A, nullptr, Yo, tile, t);
    }
  };
  auto pass1_inv = [&](const float2* Yi, float2* out) {
    for (int tile = 0; tile < P1::TILES; ++tile) {
      float2* A = reinterpret_cast<float2*>(smem1.data());
      unsigned short* T1 = P1::table(A);
      for (int t = 0; t < THREADS; ++t) P1::tables(T1, t);
      // the phases of inverse_from_rows
      for (int t = 0; t < THREADS; ++t) P1::load_rows(Yi, A, tile, t);
      for (int t = 0; t < THREADS; ++t) P1::template stage_n<true, 0, 0>(A, nullptr, nullptr, tile, t);
      for (int t = 0; t < THREADS; ++t) P1::template stage_n<true, 1, 0>(A, nullptr, nullptr, tile, t);
      for (int t = 0; t < THREADS; ++t) P1::template stage_n<true, 2, 0>(A, nullptr, nullptr, tile, t);
      for (int t = 0; t < THREADS; ++t) P1::store_natural(A, out, T1, tile, t);
    }
  };
  auto stages2 = [&](float2* A, int tile, bool inv) {
    if (!inv) {
      for (int t = 0; t < THREADS; ++t) P2::template stage_d<false>(A, tile, t);
      for (int t = 0; t < THREADS; ++t) P2::template stage_e<false>(A, tile, t);
      for (int t = 0; t < THREADS; ++t) P2::template stage_f<false>(A, tile, t);
    } else {
      for (int t = 0; t < THREADS; ++t) P2::template stage_d<true>(A, tile, t);
      for (int t = 0; t < THREADS; ++t) P2::template stage_e<true>(A, tile, t);
      for (int t = 0; t < THREADS; ++t) P2::template stage_f<true>(A, tile, t);
    }
  };
  // ---- rfft: x -> X * scale
  pass1_fwd(reinterpret_cast<const float2*>(x.data()), Y.data());
  for (int tile = 0; tile < P2::TILES; ++tile) {
    float2* A = reinterpret_cast<float2*>(smem2.data());
    for (int t = 0; t < THREADS; ++t) P2::tables(A, twls.data(), t);
    for (int t = 0; t < THREADS; ++t) P2::load_rows(Y.data(), A, tile, t);
    stages2(A, tile, false);
    for (int t = 0; t < THREADS; ++t) P2::post_to_x(A, X.data(), twls.data(), scale.data(), tile, t);
  }
  // ---- spectral filter: y = irfft(rfft(x) H)
  for (int tile = 0; tile < P2::TILES; ++tile) {
    float2* A = reinterpret_cast<float2*>(smem2.data());
    for (int t = 0; t < THREADS; ++t) P2::tables(A, twls.data(), t);
    for (int t = 0; t < THREADS; ++t) P2::load_rows(Y.data(), A, tile, t);
    stages2(A, tile, false);
    for (int t = 0; t < THREADS; ++t) P2::mid_filter(A, twls.data(), Hf.data(), tile, t);
    stages2(A, tile, true);
    for (int t = 0; t < THREADS; ++t) P2::store_rows(A, Y2.data(), tile, t);
  }
  pass1_inv(Y2.data(), y.data());
  // ---- irfft(X * scale2) with scale2 = H
  GatherTab none{nullptr, nullptr};
  for (int tile = 0; tile < P2::TILES; ++tile) {
    float2* A = reinterpret_cast<float2*>(smem2.data());
    for (int t = 0; t < THREADS; ++t) P2::tables(A, twls.data(), t);
    for (int t = 0; t < THREADS; ++t)
      P2::template pre_from_x<false>(A, X.data(), none, twls.data(), Hf.data(), tile, t);
    stages2(A, tile, true);
    for (int t = 0; t < THREADS; ++t) P2::store_rows(A, Y2.data(), tile, t);
  }
  pass1_inv(Y2.data(), xr.data());
  // ---- gather: every bin is the sum of 0..3 entries of a pool
  const int pool = 3 * (Nc + 1);
  std::vector<float2> BS(pool + 1);
  for (auto& v : BS) v = make_float2(rnd(), rnd());
  BS[pool] = make_float2(0.f, 0.f);                     // the "no band" entry
  std::vector<int4> src(Nc + 1);
  for (int k = 0; k <= Nc; ++k) {
    int s[4];
    const int cnt = (k * 7 + 3) % 4;
    for (int q = 0; q < 4; ++q) s[q] = q < cnt ? (int)(((long long)k * 2654435761LL + q * 40503) % pool) : pool;
    src[k] = make_int4(s[0], s[1], s[2], s[3]);
  }
  GatherTab g{BS.data(), src.data()};
  for (int tile = 0; tile < P2::TILES; ++tile) {
    float2* A = reinterpret_cast<float2*>(smem2.data());
    for (int t = 0; t < THREADS; ++t) P2::tables(A, twls.data(), t);
    for (int t = 0; t < THREADS; ++t)
      P2::template pre_from_x<true>(A, nullptr, g, twls.data(), scale.data(), tile, t);
    stages2(A, tile, true);
    for (int t = 0; t < THREADS; ++t) P2::store_rows(A, Y2.data(), tile, t);
  }
  pass1_inv(Y2.data(), xg.data());
  dump(dir, "x.f32", x.data(), x.size() * 4);
  dump(dir, "H.f32", Hf.data(), Hf.size() * 4);
  dump(dir, "scale.f32", scale.data(), scale.size() * 4);
  dump(dir, "X.c64", X.data(), X.size() * 8);
  dump(dir, "y.f32", y.data(), y.size() * 8);
  dump(dir, "xr.f32", xr.data(), xr.size() * 8);
  dump(dir, "BS.c64", BS.data(), BS.size() * 8);
  dump(dir, "src.i32", src.data(), src.size() * 16);
  dump(dir, "xg.f32", xg.data(), xg.size() * 8);
  printf("ok Nc=%d N1=%d N2=%d\n", Nc, PL::N1, PL::N2);
}

int main(int argc, char** argv) {
  if (argc < 3) return 1;
  const int plan = atoi(argv[1]);
  const std::string dir = argv[2];
  if (plan == 0) run<Plan<4, 3, 1, 5, 7, 1>, 16>(dir);              // Nc = 420 (even N1, two-digit passes)
  else if (plan == 1) run<Plan<3, 5, 1, 4, 7, 1>, 16>(dir);         // odd N1: no self-mirrored k1 = N1 / 2
  else if (plan == 2) run<Plan<4, 7, 11, 13, 23, 1>, 16>(dir);      // Ls = 184184 (BASELINE configs[1])
  else if (plan == 3) run<Plan<8, 7, 11, 13, 23, 1>, 16>(dir);      // Ls = 368368 (44.1 kHz, 8.35 s)
  else if (plan == 4) run<Plan<4, 7, 11, 13, 23, 1>, 8>(dir);       // 8-column tiles (small batches)
  else if (plan == 5) run<Plan<4, 3, 1, 5, 7, 1>, 8>(dir);
  else if (plan == 6) run<Plan<27, 25, 1, 49, 2, 1>, 16>(dir);      // Ls = 132300 (prime powers 27, 25, 49 as single digits)
  else if (plan == 7) run<Plan<25, 9, 2, 49, 11, 1>, 8>(dir);       // Ls = 485100
  else return 1;
  return 0;
}
