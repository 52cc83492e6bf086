// Parametric piecewise dB-per-octave lowpass, device side.
//
// Restates the semantics of design_filter (utils/blind_bwe_utils.py:82-119 of
// eloimoliner/BABE) without the sequential masked overwrites: bin k is owned by
// the LAST breakpoint i whose first bin kf[i] (first k with f[k] >= fc[i]) is
// <= k, and the gain of segment i is anchored at the value its predecessor
// took at bin kf[i].  The fp32 operation order of the reference
// (A*log2(f/fc) -> /20 -> 10** -> *anchor) is kept.
#pragma once
#include "common.cuh"

namespace babe {

struct FilterSegs {
  int K;
  int kf[BABE_MAX_BREAKPOINTS];       // first bin >= fc[i] (F if none)
  int parent[BABE_MAX_BREAKPOINTS];   // owner of bin kf[i] before segment i is written (-1: H=1)
  float fc[BABE_MAX_BREAKPOINTS];
  float A[BABE_MAX_BREAKPOINTS];
  float anchor[BABE_MAX_BREAKPOINTS]; // multiplicative anchor of segment i
  int bad;                            // some fc[i>=1] above the last bin (reference: IndexError)
};

// Round-to-nearest fp32 division / product that the compiler may not contract or approximate; on the
// host (tests/host/filter_design_host_check.cu) plain IEEE operations through volatiles.
BABE_HD float rn_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdiv_rn(a, b);
#else
  volatile float q = a / b;
  return q;
#endif
}
BABE_HD float rn_mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  volatile float q = a * b;
  return q;
#endif
}

BABE_HD float rn_sub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  volatile float q = a - b;
  return q;
#endif
}

BABE_HD float seg_gain(float A, float fc, float f) {
  // 10 ** (A * log2(f / fc) / 20)
  const float t = rn_div(rn_mul(A, log2f(rn_div(f, fc))), 20.0f);
  return exp10f(t);
}

BABE_HD int first_bin_ge(const float* f, int F, float v) {
  // f is non-decreasing; NaN v -> F
  int lo = 0, hi = F;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (f[mid] >= v) hi = mid; else lo = mid + 1;
  }
  return (v == v) ? lo : F;
}

// Executed by ONE thread.  f may live in shared or global memory.
__host__ __device__ inline void build_segments(FilterSegs& s, const float* fc, const float* A, int K,
                                      const float* f, int F) {
  s.K = K;
  s.bad = 0;
  for (int i = 0; i < K; ++i) {
    s.fc[i] = fc[i];
    s.A[i] = A[i];
    s.kf[i] = first_bin_ge(f, F, s.fc[i]);
  }
  for (int i = 0; i < K; ++i) {
    int p = -1;
    for (int j = 0; j < i; ++j)
      if (s.kf[j] <= s.kf[i]) p = j;
    s.parent[i] = p;
    if (i > 0 && s.kf[i] >= F) { s.bad = 1; s.anchor[i] = 1.0f; continue; }
    if (i == 0 || p < 0) {
      s.anchor[i] = 1.0f;
    } else {
      s.anchor[i] = rn_mul(seg_gain(s.A[p], s.fc[p], f[s.kf[i]]), s.anchor[p]);
    }
  }
}

// CTA-cooperative variant: K threads search their breakpoint in parallel (the
// binary searches are chains of dependent global loads), thread 0 then links the
// anchors from registers/shared memory only.  Contains two CTA barriers.
__device__ inline void build_segments_coop(FilterSegs& s, float* fkf /*shared, K floats*/,
                                           const float* fc, const float* A, int K, const float* f,
                                           int F) {
  const int i = threadIdx.x;
  if (i < K) {
    const float c = fc[i];
    s.fc[i] = c;
    s.A[i] = A[i];
    const int k = first_bin_ge(f, F, c);
    s.kf[i] = k;
    fkf[i] = (k < F) ? f[k] : 0.f;
  }
  __syncthreads();
  if (i == 0) {
    s.K = K;
    s.bad = 0;
    for (int a = 0; a < K; ++a) {
      int p = -1;
      for (int j = 0; j < a; ++j)
        if (s.kf[j] <= s.kf[a]) p = j;
      s.parent[a] = p;
      if (a > 0 && s.kf[a] >= F) { s.bad = 1; s.anchor[a] = 1.0f; continue; }
      s.anchor[a] = (a == 0 || p < 0) ? 1.0f
                                      : __fmul_rn(seg_gain(s.A[p], s.fc[p], fkf[a]), s.anchor[p]);
    }
  }
  __syncthreads();
}

BABE_HD int bin_owner(const FilterSegs& s, int k) {
  int o = -1;
#pragma unroll 4
  for (int i = 0; i < s.K; ++i)
    if (s.kf[i] <= k) o = i;
  return o;
}

BABE_HD float bin_gain(const FilterSegs& s, int k, float fk) {
  const int o = bin_owner(s, k);
  if (o < 0) return 1.0f;
  const float g = seg_gain(s.A[o], s.fc[o], fk);
  return (o == 0) ? g : rn_mul(g, s.anchor[o]);
}

// Chain rule of the design through the anchors.  Inputs: per-segment sums
//   s[i] = sum of u_k over the bins owned by i,  l[i] = sum of u_k log2(f_k / fc_i),  u_k = dL/dH_k * H_k.
// Executed by ONE thread.
BABE_HD void finish_param_grads(const FilterSegs& sg, const float* f, int F,
                                                   const double* s, const double* l,
                                                   double* gfc, double* gA) {
  const double alpha = 0.11512925464970229;   // ln(10)/20
  const double ln2 = 0.6931471805599453;
  for (int i = 0; i < sg.K; ++i) { gfc[i] = 0.0; gA[i] = 0.0; }
  for (int i = 0; i < sg.K; ++i) {
    if (sg.kf[i] >= F) continue;                 // owns no bin, s[i] = l[i] = 0
    gA[i] += alpha * l[i];
    gfc[i] += -alpha * (double)sg.A[i] / ((double)sg.fc[i] * ln2) * s[i];
    int c = i;
    while (sg.parent[c] >= 0) {
      const int j = sg.parent[c];
      gA[j] += alpha * (double)log2f(rn_div(f[sg.kf[c]], sg.fc[j])) * s[i];
      gfc[j] += -alpha * (double)sg.A[j] / ((double)sg.fc[j] * ln2) * s[i];
      c = j;
    }
  }
}

// Projection step of the filter fit (testing/blind_bwe_sampler.py:576-583): SEQUENTIAL clamps -- every
// breakpoint at least 1 Hz above its predecessor, every slope at most its predecessor's (or Amax).
BABE_HD void fit_project(float* fc, float* A, int K, const babe_fit_config& cfg) {
  if (cfg.clamp_fc) {
    fc[0] = fminf(fmaxf(fc[0], cfg.fcmin), cfg.fcmax);
    for (int k = 1; k < K; ++k)
#ifdef __CUDA_ARCH__
      fc[k] = fminf(fmaxf(fc[k], __fadd_rn(fc[k - 1], 1.0f)), cfg.fcmax);
#else
      fc[k] = fminf(fmaxf(fc[k], fc[k - 1] + 1.0f), cfg.fcmax);
#endif
  }
  if (cfg.clamp_A) {
    const float top0 = cfg.only_negative_A ? -1.0f : cfg.Amax;
    A[0] = fminf(fmaxf(A[0], cfg.Amin), top0);
    for (int k = 1; k < K; ++k) {
      const float top = cfg.only_negative_A ? A[k - 1] : cfg.Amax;
      A[k] = fminf(fmaxf(A[k], cfg.Amin), top);
    }
  }
}

// Stopping test (:586-588): mean absolute change of fc and of A below their tolerances.
BABE_HD bool fit_converged(const float* fc, const float* A, const float* fc_prev, const float* A_prev, int K,
                           const babe_fit_config& cfg) {
  float d0 = 0.f, d1 = 0.f;
  for (int k = 0; k < K; ++k) {
    d0 += fabsf(fc[k] - fc_prev[k]);
    d1 += fabsf(A[k] - A_prev[k]);
  }
  return d0 / (float)K < cfg.tol_fc && d1 / (float)K < cfg.tol_A;
}

}  // namespace babe
