// Real FFT of length Ls through a complex FFT of length Nc = Ls/2: the per-bin-pair post-processing
// (r2c) and pre-processing (c2r) used by k_rfft_post / k_irfft_pre / k_spectral_mid / k_cqt_gather_pre.
// __host__ __device__: tests/host/rfft_pairs_host_check.cu runs them on the CPU.
#pragma once
#include "smemfft.cuh"

namespace babe {

// Z = FFT_Nc(x_even + i x_odd)  ->  X[k], X[Nc-k]   (0 < k <= Nc/2)
BABE_HD void post_pair(float2 zk, float2 zkp, float2 W, float2& Xk, float2& Xkp) {
  const float2 a = zk, b = cconj(zkp);
  const float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
  const float2 D = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y));
  const float2 O = make_float2(D.y, -D.x);          // -i D
  const float2 T = cmul(W, O);
  Xk = make_float2(E.x + T.x, E.y + T.y);
  Xkp = make_float2(E.x - T.x, -(E.y - T.y));
}
// X[k], X[Nc-k] (Hermitian half spectrum)  ->  conj(Z[k])/Nc, conj(Z[Nc-k])/Nc
BABE_HD void pre_pair(float2 Xk, float2 Xkp, float2 W, float inv_nc,
                                         float2& Zk, float2& Zkp) {
  const float2 a = Xk, b = cconj(Xkp);
  const float2 E = make_float2(0.5f * (a.x + b.x), 0.5f * (a.y + b.y));
  const float2 D = make_float2(0.5f * (a.x - b.x), 0.5f * (a.y - b.y));
  const float2 O = cmul(cconj(W), D);
  // Z[k] = E + i O ; Z[kp] = conj(E) + i conj(O); stored conjugated and scaled
  Zk = make_float2((E.x - O.y) * inv_nc, -(E.y + O.x) * inv_nc);
  Zkp = make_float2((E.x + O.y) * inv_nc, -(-E.y + O.x) * inv_nc);
}

}  // namespace babe
