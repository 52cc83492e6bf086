// Length-Ls real FFT / inverse of the NSGT (Ls = 184184 = 2^3*7*11*13*23 in the shipped configuration) as a
// complex FFT of Nc = Ls/2 = N1*N2 points in two passes over tiles of 16 sequences, second generation.
//
// Round 1 (k_fft_cols / k_fft_rows / k_rfft_post / k_irfft_pre / k_spectral_mid / k_cqt_gather_pre) flattened
// (sequence, butterfly) tasks over the threads: 2/3 of its issue slots were index arithmetic (FastDiv, skewed
// addresses), half of its shared-memory wavefronts bank conflicts, and the r2c / c2r bin-pair processing and the
// overlap-add gather were separate passes over the spectrum (profiles/r01_cqt_B64.md).  Here
//  * a tile is held [element][sequence] (pitch 17 float2): a half-warp owns ONE butterfly index and its 16 lanes
//    are the 16 sequences, so every butterfly address / twiddle is uniform across the half-warp (one integer
//    computation per 16 butterflies, twiddles are broadcast loads) and all accesses are conflict-free;
//  * the forward transform is  K1: columns (length N1, twiddle W_Nc^{n2 k1})  ->  K2: rows (length N2) with the
//    r2c post-processing in its epilogue: a tile holds the sequence pairs (k1, N1 - k1), i.e. both members of
//    every bin pair (k, Nc - k), so X is written directly;
//  * the inverse runs the decomposition with the roles swapped (k = k1 + N1 k2 first over k2): its first pass
//    takes the bin pairs straight from the half spectrum -- or gathers them from the synthesis band spectra --
//    in its prologue (c2r pre-processing), so no pre-pass exists either.
#pragma once
#include "common.cuh"
#include "rfft_pairs.cuh"
#include "smemfft.cuh"

namespace babe {

constexpr int TF_THREADS = 256;
constexpr int TF_TW_LO = 1024;       // low part of the two-level twiddle tables
// SEQ sequences per tile (16: a half-warp per butterfly, 128-byte global segments, 128 registers -> 2 CTAs per SM;
// 8: a quarter-warp per butterfly, 64-byte segments, 80 registers -> 3 CTAs per SM), row pitch SEQ + 1 float2
template <int SEQ> struct TileCfg {
  static constexpr int PITCH = SEQ + 1, SLOTS = TF_THREADS / SEQ, HALF = SEQ / 2, CTAS = SEQ == 16 ? 2 : 3;
};

__device__ __forceinline__ float2 tf_tw2(const float2* tab, int m) {
  // exp(-2 pi i m / n) = lo[m & 1023] * hi[m >> 10]
  return cmul(__ldg(tab + (m & (TF_TW_LO - 1))), __ldg(tab + TF_TW_LO + (m >> 10)));
}

// One Stockham stage over the tile: radix R, Ns = product of the radices already applied.
template <int R, int SEQ>
__device__ __forceinline__ void tile_stage(const float2* src, float2* dst, int n, int Ns, FastDiv dns,
                                           const float2* roots, int slot, int s) {
  constexpr int TF_PITCH = TileCfg<SEQ>::PITCH;
  const int m = n / R;
  const int tw_step = m / Ns;
  for (int j = slot; j < m; j += TileCfg<SEQ>::SLOTS) {
    const int k = dns.mod(j);
    float vr[R], vi[R];
    const float2* p = src + j * TF_PITCH + s;
#pragma unroll
    for (int t = 0; t < R; ++t) {
      float2 v = p[t * m * TF_PITCH];
      if (t > 0 && k > 0) v = cmul(v, roots[t * k * tw_step]);
      vr[t] = v.x; vi[t] = v.y;
    }
    butterfly<R>(vr, vi, roots, n);
    float2* q = dst + ((j - k) * R + k) * TF_PITCH + s;
#pragma unroll
    for (int t = 0; t < R; ++t) q[t * Ns * TF_PITCH] = make_float2(vr[t], vi[t]);
  }
}

// Forward FFT of the 16 sequences of a tile; data starts in `a`, returns the buffer holding the result.
template <int SEQ>
__device__ __forceinline__ float2* tile_fft(float2* a, float2* b, const FftFactors& f, const float2* roots,
                                            int slot, int s) {
  int Ns = 1;
  float2* src = a;
  float2* dst = b;
  for (int st = 0; st < f.nf; ++st) {
    const int r = f.radix[st];
    switch (r) {
      case 2: tile_stage<2, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 3: tile_stage<3, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 4: tile_stage<4, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 5: tile_stage<5, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 7: tile_stage<7, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 8: tile_stage<8, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 11: tile_stage<11, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 13: tile_stage<13, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 16: tile_stage<16, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 17: tile_stage<17, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 19: tile_stage<19, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      case 23: tile_stage<23, SEQ>(src, dst, f.n, Ns, f.div_ns[st], roots, slot, s); break;
      default: break;
    }
    Ns *= r;
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

// ---------------------------------------------------------------------------
// F1: transforms of length N1 over columns: element e of sequence q sits at  base[e * N2 + q]  on both sides.
//   forward first pass  (TWIDDLE): out = FFT_{N1} * W_Nc^{q k1}
//   inverse last pass   (CONJ):    out = conj(FFT_{N1})           (the c2r result: x as packed complex)
// ---------------------------------------------------------------------------
struct F1Args {
  const float2* in; float2* out;
  int Nc, N1, N2;
  FftFactors f; const float2* roots; const float2* tw_nc;
  int twiddle, conj_out;
};

template <int SEQ>
__global__ void __launch_bounds__(TF_THREADS, TileCfg<SEQ>::CTAS) k_fft_n1(const F1Args a) {
  constexpr int TF_PITCH = TileCfg<SEQ>::PITCH, TF_SEQ = SEQ, SLOTS = TileCfg<SEQ>::SLOTS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* A = reinterpret_cast<float2*>(smem_raw);
  float2* Bf = A + a.N1 * TF_PITCH;
  float2* roots = Bf + a.N1 * TF_PITCH;
  const int tid = threadIdx.x, slot = tid / SEQ, s = tid % SEQ;
  for (int i = tid; i < a.N1; i += TF_THREADS) roots[i] = a.roots[i];
  const int q = blockIdx.x * TF_SEQ + s;
  const bool live = q < a.N2;
  const float2* in = a.in + (size_t)blockIdx.y * a.Nc;
  float2* out = a.out + (size_t)blockIdx.y * a.Nc;
  for (int e = slot; e < a.N1; e += SLOTS)
    A[e * TF_PITCH + s] = live ? in[(size_t)e * a.N2 + q] : make_float2(0.f, 0.f);
  __syncthreads();
  const float2* res = tile_fft<SEQ>(A, Bf, a.f, roots, slot, s);
  if (!live) return;
  for (int e = slot; e < a.N1; e += SLOTS) {
    float2 v = res[e * TF_PITCH + s];
    if (a.twiddle) v = cmul(v, tf_tw2(a.tw_nc, q * e));
    if (a.conj_out) v.y = -v.y;
    out[(size_t)e * a.N2 + q] = v;
  }
}

// ---------------------------------------------------------------------------
// F2: transforms of length N2 of the sequences k1 (rows of the [N1][N2] intermediate Y[k1 * N2 + j]); a tile holds
// 8 sequence pairs (k1, N1 - k1) -- slots 0..7 the low, 8..15 the high members -- and one extra tile the
// self-paired sequences 0 and (N1 even) N1/2, so that both bins of every pair (k, Nc - k), k = k1 + N1 k2, are in
// the tile.
// ---------------------------------------------------------------------------
struct F2Args {
  int Nc, N1, N2;
  FftFactors f; const float2* roots; const float2* tw_nc; const float2* tw_ls;
  const float* scale;          // per-bin real scale [Nc+1] (optional)
  const float2* Y; float2* Yout;   // the [N1][N2] intermediate (read by fwd, written by inv)
  const float2* X; float2* Xout;   // half spectrum [Nc+1] (written by fwd, read by inv)
  // inverse from the synthesis band spectra: X[k] = sum of <= 3 covering bands (deterministic gather)
  const float2* BS; int sum_lg;
  const int* band_p; const int* band_lg; const int* band_off; const int* jlo; const int* jhi;
};

// sequence of slot s in tile t (-1: empty slot); H = SEQ / 2 pairs per tile
template <int H>
__device__ __forceinline__ int f2_seq(int N1, int tile, int s) {
  const int P = (N1 - 1) / 2;                    // pairs (k1, N1 - k1), k1 = 1..P
  const int ntp = (P + H - 1) / H;
  if (tile < ntp) {
    const int lo = 1 + H * tile + (s % H);
    if (lo > P) return -1;
    return s < H ? lo : N1 - lo;
  }
  if (s == 0) return 0;
  if (s == 1 && (N1 & 1) == 0) return N1 / 2;
  return -1;
}

// forward second pass + r2c post-processing:  Y -> X (half spectrum, Nc+1 bins, optional per-bin scale)
template <int SEQ>
__global__ void __launch_bounds__(TF_THREADS, TileCfg<SEQ>::CTAS) k_fft_n2_fwd(const F2Args a) {
  constexpr int TF_PITCH = TileCfg<SEQ>::PITCH, TF_SEQ = SEQ, H = TileCfg<SEQ>::HALF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* A = reinterpret_cast<float2*>(smem_raw);
  float2* Bf = A + a.N2 * TF_PITCH;
  float2* roots = Bf + a.N2 * TF_PITCH;
  const int tid = threadIdx.x, slot = tid / SEQ, s = tid % SEQ;
  for (int i = tid; i < a.N2; i += TF_THREADS) roots[i] = a.roots[i];
  const int tile = blockIdx.x;
  const float2* Y = a.Y + (size_t)blockIdx.y * a.Nc;
  float2* X = a.Xout + (size_t)blockIdx.y * (a.Nc + 1);
  // rows are contiguous: lanes run along the element index, the tile is written transposed (pitch 17: conflict-free)
#pragma unroll 4
  for (int sq = 0; sq < TF_SEQ; ++sq) {
    const int k1 = f2_seq<H>(a.N1, tile, sq);
    for (int e = tid; e < a.N2; e += TF_THREADS)
      A[e * TF_PITCH + sq] = k1 >= 0 ? Y[(size_t)k1 * a.N2 + e] : make_float2(0.f, 0.f);
  }
  __syncthreads();
  const float2* res = tile_fft<SEQ>(A, Bf, a.f, roots, slot, s);
  const int ntp = ((a.N1 - 1) / 2 + H - 1) / H;
  if (tile < ntp) {
    // pair (slot i, slot i + H): Z[k1 + N1 k2] with Z[(N1 - k1) + N1 (N2 - 1 - k2)] = Z[Nc - k]
    for (int idx = tid; idx < H * a.N2; idx += TF_THREADS) {
      const int i = idx % H, k2 = idx / H;
      const int k1 = f2_seq<H>(a.N1, tile, i);
      if (k1 < 0) continue;
      const int k = k1 + a.N1 * k2, kp = a.Nc - k;
      float2 Xk, Xkp;
      post_pair(res[k2 * TF_PITCH + i], res[(a.N2 - 1 - k2) * TF_PITCH + i + H], tf_tw2(a.tw_ls, k), Xk, Xkp);
      if (a.scale) { const float sk = a.scale[k], sp = a.scale[kp]; Xk.x *= sk; Xk.y *= sk; Xkp.x *= sp; Xkp.y *= sp; }
      X[k] = Xk;
      X[kp] = Xkp;
    }
  } else {
    // sequence 0 pairs with itself (k2 <-> N2 - k2; k2 = 0 is the DC / Nyquist pair), sequence N1/2 likewise
    // (k2 <-> N2 - 1 - k2)
    for (int idx = tid; idx < 2 * a.N2; idx += TF_THREADS) {
      const int which = idx & 1, k2 = idx >> 1;
      if (which == 0) {
        if (k2 == 0) {
          const float2 z0 = res[0];
          const float s0 = a.scale ? a.scale[0] : 1.f, sn = a.scale ? a.scale[a.Nc] : 1.f;
          X[0] = make_float2((z0.x + z0.y) * s0, 0.f);
          X[a.Nc] = make_float2((z0.x - z0.y) * sn, 0.f);
          continue;
        }
        const int k2p = a.N2 - k2;
        if (k2 > k2p) continue;
        const int k = a.N1 * k2, kp = a.Nc - k;
        float2 Xk, Xkp;
        post_pair(res[k2 * TF_PITCH], res[k2p * TF_PITCH], tf_tw2(a.tw_ls, k), Xk, Xkp);
        if (a.scale) { const float sk = a.scale[k], sp = a.scale[kp]; Xk.x *= sk; Xk.y *= sk; Xkp.x *= sp; Xkp.y *= sp; }
        X[k] = Xk;
        if (kp != k) X[kp] = Xkp;
      } else if ((a.N1 & 1) == 0) {
        const int k2p = a.N2 - 1 - k2;
        if (k2 > k2p) continue;
        const int k = a.N1 / 2 + a.N1 * k2, kp = a.Nc - k;
        float2 Xk, Xkp;
        post_pair(res[k2 * TF_PITCH + 1], res[k2p * TF_PITCH + 1], tf_tw2(a.tw_ls, k), Xk, Xkp);
        if (a.scale) { const float sk = a.scale[k], sp = a.scale[kp]; Xk.x *= sk; Xk.y *= sk; Xkp.x *= sp; Xkp.y *= sp; }
        X[k] = Xk;
        if (kp != k) X[kp] = Xkp;
      }
    }
  }
}

// bin k of the half spectrum the inverse starts from: either X[k] (times an optional scale) or the overlap-add of
// the synthesis band spectra (each bin sums its <= 3 covering bands in a fixed order)
template <bool GATHER>
__device__ __forceinline__ float2 f2_bin(const F2Args& a, const float2* src, int k) {
  float2 v;
  if (!GATHER) {
    v = src[k];
  } else {
    v = make_float2(0.f, 0.f);
    const int hi = __ldg(a.jhi + k);
    for (int j = __ldg(a.jlo + k); j <= hi; ++j) {
      const int lg = __ldg(a.band_lg + j);
      const int i = k - __ldg(a.band_p + j) + lg / 2;
      if (i >= 0 && i < lg) {
        const float2 b = src[__ldg(a.band_off + j) + i];
        v.x += b.x; v.y += b.y;
      }
    }
  }
  if (a.scale) { const float sc = a.scale[k]; v.x *= sc; v.y *= sc; }
  return v;
}

// inverse first pass: c2r pre-processing of the bin pairs in the prologue (from X or gathered from the band
// spectra), FFT_{N2} over k2, twiddle W_Nc^{k1 m1}  ->  Yout[k1 * N2 + m1]
template <bool GATHER, int SEQ>
__global__ void __launch_bounds__(TF_THREADS, TileCfg<SEQ>::CTAS) k_fft_n2_inv(const F2Args a) {
  constexpr int TF_PITCH = TileCfg<SEQ>::PITCH, TF_SEQ = SEQ, H = TileCfg<SEQ>::HALF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* A = reinterpret_cast<float2*>(smem_raw);
  float2* Bf = A + a.N2 * TF_PITCH;
  float2* roots = Bf + a.N2 * TF_PITCH;
  const int tid = threadIdx.x, slot = tid / SEQ, s = tid % SEQ;
  for (int i = tid; i < a.N2; i += TF_THREADS) roots[i] = a.roots[i];
  const int tile = blockIdx.x;
  const float2* src = GATHER ? a.BS + (size_t)blockIdx.y * a.sum_lg : a.X + (size_t)blockIdx.y * (a.Nc + 1);
  float2* Y = a.Yout + (size_t)blockIdx.y * a.Nc;
  const float inv_nc = 1.0f / (float)a.Nc;
  const int ntp = ((a.N1 - 1) / 2 + H - 1) / H;
  if (tile < ntp) {
    for (int idx = tid; idx < H * a.N2; idx += TF_THREADS) {
      const int i = idx % H, k2 = idx / H;
      const int k1 = f2_seq<H>(a.N1, tile, i);
      float2 Zk = make_float2(0.f, 0.f), Zkp = Zk;
      if (k1 >= 0) {
        const int k = k1 + a.N1 * k2;
        pre_pair(f2_bin<GATHER>(a, src, k), f2_bin<GATHER>(a, src, a.Nc - k), tf_tw2(a.tw_ls, k), inv_nc, Zk, Zkp);
      }
      A[k2 * TF_PITCH + i] = Zk;
      A[(a.N2 - 1 - k2) * TF_PITCH + i + H] = Zkp;
    }
  } else {
    for (int idx = tid; idx < TF_SEQ * a.N2; idx += TF_THREADS)       // empty slots
      if ((idx % SEQ) >= 2 || ((idx % SEQ) == 1 && (a.N1 & 1))) A[(idx / SEQ) * TF_PITCH + (idx % SEQ)] = make_float2(0.f, 0.f);
    for (int idx = tid; idx < 2 * a.N2; idx += TF_THREADS) {
      const int which = idx & 1, k2 = idx >> 1;
      if (which == 0) {
        const int k2p = k2 == 0 ? 0 : a.N2 - k2;
        if (k2 > k2p && k2 != 0) continue;
        const int k = a.N1 * k2, kp = a.Nc - k;                // k2 = 0: the pair (DC, Nyquist)
        float2 xa = f2_bin<GATHER>(a, src, k), xb = f2_bin<GATHER>(a, src, kp);
        if (k2 == 0) { xa.y = 0.f; xb.y = 0.f; }
        float2 Zk, Zkp;
        pre_pair(xa, xb, tf_tw2(a.tw_ls, k), inv_nc, Zk, Zkp);
        A[k2 * TF_PITCH] = Zk;
        if (k2 != 0 && k2p != k2) A[k2p * TF_PITCH] = Zkp;
      } else if ((a.N1 & 1) == 0) {
        const int k2p = a.N2 - 1 - k2;
        if (k2 > k2p) continue;
        const int k = a.N1 / 2 + a.N1 * k2;
        float2 Zk, Zkp;
        pre_pair(f2_bin<GATHER>(a, src, k), f2_bin<GATHER>(a, src, a.Nc - k), tf_tw2(a.tw_ls, k), inv_nc, Zk, Zkp);
        A[k2 * TF_PITCH + 1] = Zk;
        if (k2p != k2) A[k2p * TF_PITCH + 1] = Zkp;
      }
    }
  }
  __syncthreads();
  const float2* res = tile_fft<SEQ>(A, Bf, a.f, roots, slot, s);
  // rows of Yout are contiguous: lanes along the element index
#pragma unroll 4
  for (int sq = 0; sq < TF_SEQ; ++sq) {
    const int k1 = f2_seq<H>(a.N1, tile, sq);
    if (k1 < 0) continue;
    for (int e = tid; e < a.N2; e += TF_THREADS)
      Y[(size_t)k1 * a.N2 + e] = cmul(res[e * TF_PITCH + sq], tf_tw2(a.tw_nc, k1 * e));
  }
}

static inline size_t tile_fft_smem(int n, int seq = 16) { return sizeof(float2) * ((size_t)2 * n * (seq + 1) + n); }

}  // namespace babe
