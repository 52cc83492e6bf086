// Element-wise glue of the CQTDiff+ residual layers around the (PyTorch / cuDNN)
// convolutions, one HBM pass each instead of ~8 forward + ~16 backward PyTorch
// kernels per layer (the caller on either side of the CQT):
//
//   layer (networks/cqtdiff+.py:470-482):
//     h = gelu( BiasFreeGroupNorm(x) * (affine(sigma) + 1) )          k_gn_stats, k_gn_film_gelu
//     y = (x + conv(h) * gate(sigma)) / sqrt(2)                       k_gate_residual
//   BiasFreeGroupNorm (networks/cqtdiff+.py:137-163): x / (std_g(x) + eps) * gamma_c with the
//   unbiased standard deviation of each (sample, group) -- no mean removal of x itself.
//
// Backward wrt the activations (parameters are frozen while sampling):
//   g_v = g_y * gate / sqrt(2)                                        k_gate_residual (x0 = null)
//   g_x = g_y / sqrt(2) + g_u s - G_r r^2 (x - mean) / ((cnt - 1) std)   k_gn_bwd_reduce, k_gn_bwd
//   with u = x s, s = r gamma (aff + 1), r = 1 / (std + eps), g_u = g_h gelu'(u),
//   G_r = sum_group g_u x gamma (aff + 1).
//
// Layout: contiguous NCHW float32; a "plane" is one (n, c) slice of P = F*T elements, a group
// is gc = C/G consecutive planes.  All reductions run in a fixed order (bitwise reproducible).
#include <math.h>

#include <algorithm>

#include "common.cuh"

namespace babe {

constexpr int NET_THREADS = 256;
constexpr int MAX_STAT_SLICES = 64;

__device__ __forceinline__ float gelu_f(float u) {
  return 0.5f * u * (1.0f + erff(u * 0.70710678118654752f));
}
__device__ __forceinline__ float gelu_grad_f(float u) {
  const float cdf = 0.5f * (1.0f + erff(u * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * u * u);
  return cdf + u * pdf;
}

__device__ __forceinline__ double block_sum_d(double v, double* scratch /*NET_THREADS/32*/) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double s = 0.0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < NET_THREADS / 32; ++w) s += scratch[w];
  }
  return s;          // valid in thread 0
}

// (sum, sum of squares) of slice blockIdx.x of group blockIdx.y -> part[group][slice][2]
__global__ void __launch_bounds__(NET_THREADS) k_gn_stats(const float* __restrict__ x,
                                                          double* __restrict__ part, long long cnt,
                                                          int S) {
  __shared__ double scratch[NET_THREADS / 32];
  const long long per = ((cnt + S - 1) / S + 3) & ~3LL;
  const long long lo = per * blockIdx.x, hi = min(cnt, lo + per);
  const float* g = x + (size_t)blockIdx.y * cnt;
  double s1 = 0.0, s2 = 0.0;
  const bool vec = ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  if (vec) {
    const long long n4 = (hi > lo) ? (hi - lo) / 4 : 0;
    const float4* g4 = reinterpret_cast<const float4*>(g + lo);
    // <= 32 values per float accumulator between two double updates
    for (long long i = threadIdx.x; i < n4; i += 8LL * NET_THREADS) {
      float a1 = 0.f, a2 = 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const long long j = i + (long long)u * NET_THREADS;
        if (j < n4) {
          const float4 v = g4[j];
          a1 += (v.x + v.y) + (v.z + v.w);
          a2 += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
      }
      s1 += (double)a1; s2 += (double)a2;
    }
    for (long long i = lo + n4 * 4 + threadIdx.x; i < hi; i += NET_THREADS) {
      const float v = g[i];
      s1 += (double)v; s2 += (double)v * (double)v;
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += NET_THREADS) {
      const float v = g[i];
      s1 += (double)v; s2 += (double)v * (double)v;
    }
  }
  s1 = block_sum_d(s1, scratch);
  s2 = block_sum_d(s2, scratch);
  if (threadIdx.x == 0) {
    double* p = part + ((size_t)blockIdx.y * S + blockIdx.x) * 2;
    p[0] = s1; p[1] = s2;
  }
}

struct GroupStats { float mean, r, std; };

// executed by warp 0 (fixed summation order); the result is valid in lane 0
__device__ __forceinline__ GroupStats finish_stats(const double* part, int S, long long cnt, float eps) {
  const int lane = threadIdx.x & 31;
  double s1 = 0.0, s2 = 0.0;
  for (int s = lane; s < S; s += 32) { s1 += part[2 * s]; s2 += part[2 * s + 1]; }
  s1 = warp_sum(s1); s2 = warp_sum(s2);
  const double mean = s1 / (double)cnt;
  double var = (s2 - (double)cnt * mean * mean) / (double)(cnt - 1);
  if (var < 0.0) var = 0.0;
  GroupStats g;
  g.std = (float)sqrt(var);
  g.mean = (float)mean;
  g.r = 1.0f / (g.std + eps);
  return g;
}

struct PlaneArgs {
  int N, C, G, S;          // S: statistic slices per group
  long long P;             // plane size F*T
  float eps, scale;
  const double* part;      // [N*G][S][2]
  const float* gamma;      // [C]
  const float* aff;        // [N][C]
  const float* gate;       // [N][C] or null (= 1)
  const double* gr_part;   // [N*C][S2]
  int S2;
};

// h = gelu(x * r * gamma * (aff + 1));  grid (chunks, N*C)
__global__ void __launch_bounds__(NET_THREADS) k_gn_film_gelu(const float* __restrict__ x,
                                                              float* __restrict__ h, const PlaneArgs a) {
  __shared__ float s_s;
  const int pc = blockIdx.y, n = pc / a.C, c = pc - n * a.C;
  const int gc = a.C / a.G;
  if (threadIdx.x < 32) {
    const GroupStats st = finish_stats(a.part + ((size_t)n * a.G + c / gc) * a.S * 2, a.S, (long long)gc * a.P, a.eps);
    if (threadIdx.x == 0) s_s = st.r * a.gamma[c] * (a.aff[pc] + 1.0f);
  }
  __syncthreads();
  const float s = s_s;
  const float* xp = x + (size_t)pc * a.P;
  float* hp = h + (size_t)pc * a.P;
  const long long chunk = 4LL * NET_THREADS * 4;
  const long long lo = chunk * blockIdx.x, hi = min(a.P, lo + chunk);
  if (((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(hp)) & 15) == 0) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = lo + 4LL * (threadIdx.x + u * NET_THREADS);
      if (i + 3 < hi) {
        float4 v = *reinterpret_cast<const float4*>(xp + i);
        v.x = gelu_f(v.x * s); v.y = gelu_f(v.y * s); v.z = gelu_f(v.z * s); v.w = gelu_f(v.w * s);
        *reinterpret_cast<float4*>(hp + i) = v;
      } else {
        for (long long j = i; j < hi; ++j) hp[j] = gelu_f(xp[j] * s);
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += NET_THREADS) hp[i] = gelu_f(xp[i] * s);
  }
}

// out = (x0 + v * gate) * scale   (x0 may be null: out = v * gate * scale); grid (chunks, N*C)
__global__ void __launch_bounds__(NET_THREADS) k_gate_residual(const float* __restrict__ x0,
                                                               const float* __restrict__ v,
                                                               float* __restrict__ out, const PlaneArgs a) {
  const int pc = blockIdx.y;
  const float gs = (a.gate ? a.gate[pc] : 1.0f) * a.scale, sc = a.scale;
  const float* vp = v + (size_t)pc * a.P;
  const float* xp = x0 ? x0 + (size_t)pc * a.P : nullptr;
  float* op = out + (size_t)pc * a.P;
  const long long chunk = 4LL * NET_THREADS * 4;
  const long long lo = chunk * blockIdx.x, hi = min(a.P, lo + chunk);
  const bool vec = ((reinterpret_cast<uintptr_t>(vp) | reinterpret_cast<uintptr_t>(op) |
                     reinterpret_cast<uintptr_t>(xp)) & 15) == 0;
  if (vec) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = lo + 4LL * (threadIdx.x + u * NET_THREADS);
      if (i + 3 < hi) {
        float4 q = *reinterpret_cast<const float4*>(vp + i);
        if (xp) {
          const float4 b = *reinterpret_cast<const float4*>(xp + i);
          q.x = fmaf(q.x, gs, b.x * sc); q.y = fmaf(q.y, gs, b.y * sc);
          q.z = fmaf(q.z, gs, b.z * sc); q.w = fmaf(q.w, gs, b.w * sc);
        } else {
          q.x *= gs; q.y *= gs; q.z *= gs; q.w *= gs;
        }
        *reinterpret_cast<float4*>(op + i) = q;
      } else {
        for (long long j = i; j < hi; ++j) op[j] = xp ? fmaf(vp[j], gs, xp[j] * sc) : vp[j] * gs;
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += NET_THREADS)
      op[i] = xp ? fmaf(vp[i], gs, xp[i] * sc) : vp[i] * gs;
  }
}

// partial sums of g_u * x * gamma (aff + 1) over slice blockIdx.x of plane blockIdx.y
__global__ void __launch_bounds__(NET_THREADS) k_gn_bwd_reduce(const float* __restrict__ gh,
                                                               const float* __restrict__ x,
                                                               double* __restrict__ gr_part,
                                                               const PlaneArgs a) {
  __shared__ double scratch[NET_THREADS / 32];
  __shared__ float s_r;
  const int pc = blockIdx.y, n = pc / a.C, c = pc - n * a.C;
  const int gc = a.C / a.G;
  if (threadIdx.x < 32) {
    const float r = finish_stats(a.part + ((size_t)n * a.G + c / gc) * a.S * 2, a.S, (long long)gc * a.P, a.eps).r;
    if (threadIdx.x == 0) s_r = r;
  }
  __syncthreads();
  const float q = a.gamma[c] * (a.aff[pc] + 1.0f), s = s_r * q;
  const float* xp = x + (size_t)pc * a.P;
  const float* gp = gh + (size_t)pc * a.P;
  const long long per = ((a.P + a.S2 - 1) / a.S2 + 3) & ~3LL;
  const long long lo = per * blockIdx.x, hi = min(a.P, lo + per);
  double acc = 0.0;
  if (((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(gp)) & 15) == 0) {
    const long long n4 = (hi > lo) ? (hi - lo) / 4 : 0;
    const float4* x4 = reinterpret_cast<const float4*>(xp + lo);
    const float4* g4 = reinterpret_cast<const float4*>(gp + lo);
    for (long long i = threadIdx.x; i < n4; i += 4LL * NET_THREADS) {
      float f = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long j = i + (long long)u * NET_THREADS;
        if (j < n4) {
          const float4 xv = x4[j], gv = g4[j];
          f += (gv.x * gelu_grad_f(xv.x * s) * xv.x + gv.y * gelu_grad_f(xv.y * s) * xv.y) +
               (gv.z * gelu_grad_f(xv.z * s) * xv.z + gv.w * gelu_grad_f(xv.w * s) * xv.w);
        }
      }
      acc += (double)f;
    }
    for (long long i = lo + n4 * 4 + threadIdx.x; i < hi; i += NET_THREADS)
      acc += (double)(gp[i] * gelu_grad_f(xp[i] * s) * xp[i]);
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += NET_THREADS)
      acc += (double)(gp[i] * gelu_grad_f(xp[i] * s) * xp[i]);
  }
  acc = block_sum_d(acc, scratch);
  if (threadIdx.x == 0) gr_part[(size_t)pc * a.S2 + blockIdx.x] = acc * (double)q;
}

// g_x = g_y * scale + g_h gelu'(x s) s - coef (x - mean),  coef = G_r r^2 / ((cnt - 1) std)
__global__ void __launch_bounds__(NET_THREADS) k_gn_bwd(const float* __restrict__ gh,
                                                        const float* __restrict__ x,
                                                        const float* __restrict__ gy,
                                                        float* __restrict__ gx, const PlaneArgs a) {
  __shared__ float s_s, s_coef, s_mean;
  __shared__ double scratch[NET_THREADS / 32];
  const int pc = blockIdx.y, n = pc / a.C, c = pc - n * a.C;
  const int gc = a.C / a.G, g = c / gc;
  const long long cnt = (long long)gc * a.P;
  double Gr = 0.0;
  {
    const double* grp = a.gr_part + ((size_t)n * a.C + (size_t)g * gc) * a.S2;
    for (int i = threadIdx.x; i < gc * a.S2; i += NET_THREADS) Gr += grp[i];
    Gr = block_sum_d(Gr, scratch);                       // valid in thread 0
  }
  if (threadIdx.x < 32) {
    const GroupStats st = finish_stats(a.part + ((size_t)n * a.G + g) * a.S * 2, a.S, cnt, a.eps);
    if (threadIdx.x == 0) {
      s_s = st.r * a.gamma[c] * (a.aff[pc] + 1.0f);
      s_mean = st.mean;
      s_coef = (float)(Gr * (double)st.r * (double)st.r / ((double)(cnt - 1) * (double)st.std));
    }
  }
  __syncthreads();
  const float s = s_s, coef = s_coef, mean = s_mean, sc = a.scale;
  const float* xp = x + (size_t)pc * a.P;
  const float* gp = gh + (size_t)pc * a.P;
  const float* yp = gy ? gy + (size_t)pc * a.P : nullptr;
  float* op = gx + (size_t)pc * a.P;
  const long long chunk = 4LL * NET_THREADS * 4;
  const long long lo = chunk * blockIdx.x, hi = min(a.P, lo + chunk);
  auto f = [&](float xv, float gv, float yv) {
    return fmaf(gv * gelu_grad_f(xv * s), s, fmaf(-coef, xv - mean, yv * sc));
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(xp) | reinterpret_cast<uintptr_t>(gp) |
                     reinterpret_cast<uintptr_t>(yp) | reinterpret_cast<uintptr_t>(op)) & 15) == 0;
  if (vec) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long i = lo + 4LL * (threadIdx.x + u * NET_THREADS);
      if (i + 3 < hi) {
        const float4 xv = *reinterpret_cast<const float4*>(xp + i);
        const float4 gv = *reinterpret_cast<const float4*>(gp + i);
        float4 yv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yp) yv = *reinterpret_cast<const float4*>(yp + i);
        float4 o;
        o.x = f(xv.x, gv.x, yv.x); o.y = f(xv.y, gv.y, yv.y);
        o.z = f(xv.z, gv.z, yv.z); o.w = f(xv.w, gv.w, yv.w);
        *reinterpret_cast<float4*>(op + i) = o;
      } else {
        for (long long j = i; j < hi; ++j) op[j] = f(xp[j], gp[j], yp ? yp[j] : 0.f);
      }
    }
  } else {
    for (long long i = lo + threadIdx.x; i < hi; i += NET_THREADS) op[i] = f(xp[i], gp[i], yp ? yp[i] : 0.f);
  }
}


// ---------------------------------------------------------------------------
// x2 anti-aliased time resampling of UpDownResample (networks/cqtdiff+.py:522-580, mode "T"):
// every (n, c, f) row is filtered independently by the L-tap kernel (the reference builds a dense
// diagonal C x C x L weight per call, :564-570), reflect padding fused, and the two adjoints.
//   down : xp = reflect_pad(x, pad),            y[t] = sum_k w[k] xp[2 t + k]          (T -> T/2)
//   up   : xp = reflect_pad(x, (pad + 1) / 2),  y[j] = sum_k w[k] xp[(j + 2 pad + 1 - k) / 2]
//          over the k of matching parity                                              (T -> 2 T)
// with pad = L/2 - 1.  All four kernels are gathers (deterministic).
// ---------------------------------------------------------------------------
template <int L>
struct Taps { float w[L]; };

__device__ __forceinline__ int reflect_idx(int i, int T) {
  if (i < 0) i = -i;
  if (i >= T) i = 2 * (T - 1) - i;
  return i;
}

// Four outputs per thread; interior threads of the 8-tap filter use aligned vector loads:
//   corr4  : out[m] = sum_k w[k] v[2 m + 1 + k],  v = in[2 o0 - 4 .. 2 o0 + 11]       (down, grad of up)
//   interp4: out[m] from u = in[o0/2 - 2 .. o0/2 + 3], 4 taps of alternating parity   (up, grad of down)
template <int L>
__device__ __forceinline__ void corr4(const float* __restrict__ in, int o0, const Taps<L>& w, float (&o)[4]) {
  float v[16];
  const float4* p = reinterpret_cast<const float4*>(in + 2 * o0 - 4);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float4 q = p[i]; v[4 * i] = q.x; v[4 * i + 1] = q.y; v[4 * i + 2] = q.z; v[4 * i + 3] = q.w; }
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    float a = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) a = fmaf(w.w[k], v[2 * m + 1 + k], a);
    o[m] = a;
  }
}
template <int L>
__device__ __forceinline__ void interp4(const float* __restrict__ in, int o0, const Taps<L>& w, float (&o)[4]) {
  float u[6];
  const float2* p = reinterpret_cast<const float2*>(in + o0 / 2 - 2);
#pragma unroll
  for (int i = 0; i < 3; ++i) { const float2 q = p[i]; u[2 * i] = q.x; u[2 * i + 1] = q.y; }
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) {
    a0 = fmaf(w.w[2 * kk + 1], u[3 - kk], a0);
    a1 = fmaf(w.w[2 * kk], u[4 - kk], a1);
    a2 = fmaf(w.w[2 * kk + 1], u[4 - kk], a2);
    a3 = fmaf(w.w[2 * kk], u[5 - kk], a3);
  }
  o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
}
__device__ __forceinline__ void store4(float* out, int o0, int n_out, const float (&o)[4], bool vec) {
  if (vec && o0 + 3 < n_out) {
    *reinterpret_cast<float4*>(out + o0) = make_float4(o[0], o[1], o[2], o[3]);
  } else {
#pragma unroll
    for (int m = 0; m < 4; ++m)
      if (o0 + m < n_out) out[o0 + m] = o[m];
  }
}
__device__ __forceinline__ bool aligned16(const void* a, const void* b) {
  return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

template <int L>
__global__ void __launch_bounds__(NET_THREADS) k_resample_down(const float* __restrict__ x,
                                                               float* __restrict__ y, int T, int To,
                                                               const Taps<L> w, long long rows) {
  constexpr int PAD = L / 2 - 1;
  const int per = (To + 3) / 4;                          // threads per row
  const unsigned gidx = blockIdx.x * NET_THREADS + threadIdx.x;   // < 2^32 (checked by the launcher)
  const unsigned row = gidx / (unsigned)per;
  if (row >= rows) return;
  const int t0 = 4 * (int)(gidx - row * (unsigned)per);
  const float* xr = x + (size_t)row * T;
  float* yr = y + (size_t)row * To;
  const bool vec = aligned16(xr, yr);
  float o[4];
  if (L == 8 && vec && t0 >= 2 && 2 * t0 + 11 < T) {
    corr4<L>(xr, t0, w, o);
  } else {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      float acc = 0.f;
      const int q0 = 2 * (t0 + m) - PAD;
      if (t0 + m < To) {
#pragma unroll
        for (int k = 0; k < L; ++k) acc = fmaf(w.w[k], xr[reflect_idx(q0 + k, T)], acc);
      }
      o[m] = acc;
    }
  }
  store4(yr, t0, To, o, vec);
}

template <int L>
__global__ void __launch_bounds__(NET_THREADS) k_resample_up(const float* __restrict__ x,
                                                             float* __restrict__ y, int T,
                                                             const Taps<L> w, long long rows) {
  constexpr int PAD = L / 2 - 1, PU = (PAD + 1) / 2;
  const int per = (2 * T + 3) / 4;
  const unsigned gidx = blockIdx.x * NET_THREADS + threadIdx.x;   // < 2^32 (checked by the launcher)
  const unsigned row = gidx / (unsigned)per;
  if (row >= rows) return;
  const int j0 = 4 * (int)(gidx - row * (unsigned)per);
  const float* xr = x + (size_t)row * T;
  float* yr = y + (size_t)row * 2 * T;
  const bool vec = aligned16(xr, yr);
  float o[4];
  if (L == 8 && vec && j0 >= 4 && j0 / 2 + 3 < T) {
    interp4<L>(xr, j0, w, o);
  } else {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int j = j0 + m;
      float acc = 0.f;
      const int par = (j + 2 * PAD + 1) & 1;
      if (j < 2 * T) {
#pragma unroll
        for (int kk = 0; kk < L / 2; ++kk) {
          // k = par + 2 kk (same parity as j + 2 pad + 1); L/2 taps per output
          const int i = (j + 2 * PAD + 1 - par) / 2 - kk;             // index into the padded row
          const float wk = par ? w.w[2 * kk + 1] : w.w[2 * kk];
          acc = fmaf(wk, xr[reflect_idx(i - PU, T)], acc);
        }
      }
      o[m] = acc;
    }
  }
  store4(yr, j0, 2 * T, o, vec);
}

// gradient of `down` wrt x: gy[To] -> gx[T]
template <int L>
__global__ void __launch_bounds__(NET_THREADS) k_resample_down_adj(const float* __restrict__ gy,
                                                                   float* __restrict__ gx, int T, int To,
                                                                   const Taps<L> w, long long rows) {
  constexpr int PAD = L / 2 - 1;
  const int per = (T + 3) / 4;
  const unsigned gidx = blockIdx.x * NET_THREADS + threadIdx.x;   // < 2^32 (checked by the launcher)
  const unsigned row = gidx / (unsigned)per;
  if (row >= rows) return;
  const int i0 = 4 * (int)(gidx - row * (unsigned)per);
  const float* gr = gy + (size_t)row * To;
  float* xr = gx + (size_t)row * T;
  const bool vec = aligned16(gr, xr);
  float o[4];
  if (L == 8 && vec && i0 >= 4 && i0 + 8 <= T) {
    interp4<L>(gr, i0, w, o);
  } else {
    // padded-row gradient at q: sum over (t, k) with 2 t + k = q
    auto gxp = [&](int q) {
      float a = 0.f;
#pragma unroll
      for (int kk = 0; kk < L / 2; ++kk) {
        const int k = (q & 1) + 2 * kk, t = (q - k) / 2;
        const float wk = (q & 1) ? w.w[2 * kk + 1] : w.w[2 * kk];
        if (t >= 0 && t < To && q - k >= 0) a = fmaf(wk, gr[t], a);
      }
      return a;
    };
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int i = i0 + m;
      float acc = 0.f;
      if (i < T) {
        acc = gxp(i + PAD);
        if (i >= 1 && i <= PAD) acc += gxp(PAD - i);                        // left reflection
        if (i <= T - 2 && i >= T - 1 - PAD) acc += gxp(2 * (T - 1) - i + PAD);   // right reflection
      }
      o[m] = acc;
    }
  }
  store4(xr, i0, T, o, vec);
}

// gradient of `up` wrt x: gy[2T] -> gx[T]
template <int L>
__global__ void __launch_bounds__(NET_THREADS) k_resample_up_adj(const float* __restrict__ gy,
                                                                 float* __restrict__ gx, int T,
                                                                 const Taps<L> w, long long rows) {
  constexpr int PAD = L / 2 - 1, PU = (PAD + 1) / 2;
  const int per = (T + 3) / 4;
  const unsigned gidx = blockIdx.x * NET_THREADS + threadIdx.x;   // < 2^32 (checked by the launcher)
  const unsigned row = gidx / (unsigned)per;
  if (row >= rows) return;
  const int i0 = 4 * (int)(gidx - row * (unsigned)per);
  const float* gr = gy + (size_t)row * 2 * T;
  float* xr = gx + (size_t)row * T;
  const bool vec = aligned16(gr, xr);
  float o[4];
  if (L == 8 && vec && i0 >= 4 && i0 + 8 <= T) {
    corr4<L>(gr, i0, w, o);
  } else {
    // padded-row gradient at q: sum_k w[k] gy[2 q + k - (2 pad + 1)]
    auto gxp = [&](int q) {
      float a = 0.f;
      const int j0 = 2 * q - (2 * PAD + 1);
#pragma unroll
      for (int k = 0; k < L; ++k) {
        const int j = j0 + k;
        if (j >= 0 && j < 2 * T) a = fmaf(w.w[k], gr[j], a);
      }
      return a;
    };
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const int i = i0 + m;
      float acc = 0.f;
      if (i < T) {
        acc = gxp(i + PU);
        if (i >= 1 && i <= PU) acc += gxp(PU - i);
        if (i <= T - 2 && i >= T - 1 - PU) acc += gxp(2 * (T - 1) - i + PU);
      }
      o[m] = acc;
    }
  }
  store4(xr, i0, T, o, vec);
}

template <int L>
static int launch_resample(const float* x, float* y, long long rows, int T, int mode, const float* taps,
                           cudaStream_t st) {
  Taps<L> w;
  for (int k = 0; k < L; ++k) w.w[k] = taps[k];
  const int To = T / 2;
  const int n_out = mode == 0 ? To : (mode == 1 ? 2 * T : T);
  const long long threads = rows * ((n_out + 3) / 4);
  const unsigned grid = (unsigned)((threads + NET_THREADS - 1) / NET_THREADS);
  switch (mode) {
    case 0: k_resample_down<L><<<grid, NET_THREADS, 0, st>>>(x, y, T, To, w, rows); break;
    case 1: k_resample_up<L><<<grid, NET_THREADS, 0, st>>>(x, y, T, w, rows); break;
    case 2: k_resample_down_adj<L><<<grid, NET_THREADS, 0, st>>>(x, y, T, To, w, rows); break;
    default: k_resample_up_adj<L><<<grid, NET_THREADS, 0, st>>>(x, y, T, w, rows); break;
  }
  return check_launch("k_resample");
}

static int check_dims(int N, int C, int G, long long P, const char* what) {
  BABE_REQUIRE(N >= 1 && C >= 1 && G >= 1 && P >= 1 && C % G == 0, BABE_EBADARG,
               "%s: bad shape N=%d C=%d G=%d P=%lld", what, N, C, G, P);
  BABE_REQUIRE((long long)N * C <= 65535, BABE_EUNSUPPORTED, "%s: N*C = %lld exceeds 65535", what,
               (long long)N * C);
  return BABE_OK;
}
static inline unsigned chunks_of(long long P) { return (unsigned)((P + 4LL * NET_THREADS * 4 - 1) / (4LL * NET_THREADS * 4)); }

}  // namespace babe

using namespace babe;

extern "C" int babe_gn_slices(int N, int C, int G, long long P) {
  if (N < 1 || C < 1 || G < 1 || P < 1 || C % G) return 0;
  const long long cnt = (long long)(C / G) * P;
  long long s = (16LL * sm_count() + (long long)N * G - 1) / ((long long)N * G);
  s = std::min<long long>(s, (cnt + 8191) / 8192);
  return (int)std::max<long long>(1, std::min<long long>(s, MAX_STAT_SLICES));
}

extern "C" int babe_gn_stats(const float* x, double* part, int N, int C, int G, long long P,
                             int slices, void* stream) {
  int rc = check_dims(N, C, G, P, "gn_stats");
  if (rc) return rc;
  BABE_REQUIRE(x && part && slices >= 1 && slices <= MAX_STAT_SLICES, BABE_EBADARG, "gn_stats: bad arguments");
  k_gn_stats<<<dim3(slices, N * G), NET_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(
      x, part, (long long)(C / G) * P, slices);
  return check_launch("k_gn_stats");
}

extern "C" int babe_gn_film_gelu(const float* x, float* h, const double* part, int slices,
                                 const float* gamma, const float* aff, int N, int C, int G,
                                 long long P, float eps, void* stream) {
  int rc = check_dims(N, C, G, P, "gn_film_gelu");
  if (rc) return rc;
  BABE_REQUIRE(x && h && part && gamma && aff && slices >= 1, BABE_EBADARG, "gn_film_gelu: bad arguments");
  PlaneArgs a{};
  a.N = N; a.C = C; a.G = G; a.S = slices; a.P = P; a.eps = eps; a.part = part; a.gamma = gamma; a.aff = aff;
  k_gn_film_gelu<<<dim3(chunks_of(P), N * C), NET_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x, h, a);
  return check_launch("k_gn_film_gelu");
}

extern "C" int babe_gate_residual(const float* x0, const float* v, const float* gate, float* out,
                                  int N, int C, long long P, float scale, void* stream) {
  int rc = check_dims(N, C, 1, P, "gate_residual");
  if (rc) return rc;
  BABE_REQUIRE(v && out, BABE_EBADARG, "gate_residual: bad arguments");
  PlaneArgs a{};
  a.N = N; a.C = C; a.G = 1; a.P = P; a.scale = scale; a.gate = gate;
  k_gate_residual<<<dim3(chunks_of(P), N * C), NET_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(x0, v, out, a);
  return check_launch("k_gate_residual");
}

extern "C" int babe_gn_bwd_slices(int N, int C, long long P) {
  if (N < 1 || C < 1 || P < 1) return 0;
  long long s = (16LL * sm_count() + (long long)N * C - 1) / ((long long)N * C);
  s = std::min<long long>(s, (P + 4095) / 4096);
  return (int)std::max<long long>(1, std::min<long long>(s, 16));
}

extern "C" int babe_gn_film_gelu_bwd_reduce(const float* gh, const float* x, const double* part,
                                            int slices, double* gr_part, int slices2,
                                            const float* gamma, const float* aff, int N, int C, int G,
                                            long long P, float eps, void* stream) {
  int rc = check_dims(N, C, G, P, "gn_film_gelu_bwd_reduce");
  if (rc) return rc;
  BABE_REQUIRE(gh && x && part && gr_part && gamma && aff && slices >= 1 && slices2 >= 1, BABE_EBADARG,
               "gn_film_gelu_bwd_reduce: bad arguments");
  PlaneArgs a{};
  a.N = N; a.C = C; a.G = G; a.S = slices; a.P = P; a.eps = eps;
  a.part = part; a.gamma = gamma; a.aff = aff; a.gr_part = gr_part; a.S2 = slices2;
  k_gn_bwd_reduce<<<dim3(slices2, N * C), NET_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(gh, x, gr_part, a);
  return check_launch("k_gn_bwd_reduce");
}

extern "C" int babe_gn_film_gelu_bwd(const float* gh, const float* x, const float* gy, float* gx,
                                     const double* part, int slices, const double* gr_part, int slices2,
                                     const float* gamma, const float* aff, int N, int C, int G,
                                     long long P, float eps, float res_scale, void* stream) {
  int rc = check_dims(N, C, G, P, "gn_film_gelu_bwd");
  if (rc) return rc;
  BABE_REQUIRE(gh && x && gx && part && gr_part && gamma && aff && slices >= 1 && slices2 >= 1,
               BABE_EBADARG, "gn_film_gelu_bwd: bad arguments");
  PlaneArgs a{};
  a.N = N; a.C = C; a.G = G; a.S = slices; a.P = P; a.eps = eps; a.scale = res_scale;
  a.part = part; a.gamma = gamma; a.aff = aff; a.gr_part = gr_part; a.S2 = slices2;
  k_gn_bwd<<<dim3(chunks_of(P), N * C), NET_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(gh, x, gy, gx, a);
  return check_launch("k_gn_bwd");
}

extern "C" int babe_resample2(const float* in, float* out, long long rows, int T, int mode,
                              const float* taps_host, int L, void* stream) {
  BABE_REQUIRE(in && out && taps_host && rows >= 1, BABE_EBADARG, "resample2: bad arguments");
  BABE_REQUIRE(mode >= 0 && mode <= 3, BABE_EBADARG, "resample2: mode %d", mode);
  BABE_REQUIRE(T >= L && T % 2 == 0, BABE_EUNSUPPORTED, "resample2: row length %d (even, >= %d)", T, L);
  BABE_REQUIRE(rows * (long long)T < (1LL << 31), BABE_EUNSUPPORTED, "resample2: more than 2^31 samples");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc;
  switch (L) {
    case 4: rc = launch_resample<4>(in, out, rows, T, mode, taps_host, st); break;
    case 8: rc = launch_resample<8>(in, out, rows, T, mode, taps_host, st); break;
    case 12: rc = launch_resample<12>(in, out, rows, T, mode, taps_host, st); break;
    default: BABE_REQUIRE(false, BABE_EUNSUPPORTED, "resample2: %d taps (4, 8 or 12)", L);
  }
  return rc;
}
