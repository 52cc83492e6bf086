"""Oracle: STFT -> parametric lowpass -> iSTFT operator of BABE (CPU, torch).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates
``utils/blind_bwe_utils.py`` of eloimoliner/BABE with explicit frame
arithmetic (no ``torch.stft`` / ``torch.istft``) so that every step the CUDA
kernels fuse is visible, and so that the same code runs in fp64.

Notation: N = NFFT, h = N/2, w = periodic Hamming window, frames
M = 1 + T // h (the reference right-pads N zeros and uses center=False),
F = N/2 + 1 bins, L_out = N + h (M - 1).
"""
import math

import torch


# --------------------------------------------------------------------------
# framing helpers
# --------------------------------------------------------------------------
def hamming_periodic(nfft, dtype=torch.float32):
    """``torch.hamming_window(N)`` default (periodic=True, alpha=.54, beta=.46)
    as used at utils/blind_bwe_utils.py:19 and :32."""
    n = torch.arange(nfft, dtype=torch.float64)
    return (0.54 - 0.46 * torch.cos(2.0 * math.pi * n / nfft)).to(dtype)


def num_frames(T, nfft):
    """Frames produced by apply_stft: padded length T+N, hop N/2, center=False
    (utils/blind_bwe_utils.py:22-23)."""
    return 1 + T // (nfft // 2)


def ola_envelope(frames, nfft, dtype=torch.float64):
    """Sum of squared windows, i.e. the divisor torch.istft applies
    (utils/blind_bwe_utils.py:37).  Length N + h (frames - 1)."""
    h = nfft // 2
    w2 = hamming_periodic(nfft, torch.float64) ** 2
    env = torch.zeros(nfft + h * (frames - 1), dtype=torch.float64)
    for m in range(frames):
        env[m * h:m * h + nfft] += w2
    return env.to(dtype)


# --------------------------------------------------------------------------
# a1 / a2 / a3
# --------------------------------------------------------------------------
def apply_stft(x, nfft):
    """utils/blind_bwe_utils.py:15-26.  x (B,T) -> (B,F,frames,2)."""
    B, T = x.shape
    h = nfft // 2
    w = hamming_periodic(nfft, x.dtype)
    xp = torch.cat((x, torch.zeros(B, nfft, dtype=x.dtype)), 1)
    fr = xp.unfold(1, nfft, h)                      # (B, frames, N)
    X = torch.fft.rfft(fr * w, dim=-1)              # (B, frames, F)
    X = X.transpose(1, 2)                           # (B, F, frames)
    return torch.view_as_real(X.contiguous())


def apply_filter_istft(X, H, nfft):
    """utils/blind_bwe_utils.py:28-39.  X (B,F,frames,2), H (F,) ->
    (B, N + h (frames-1))."""
    B, F, M, _ = X.shape
    h = nfft // 2
    w = hamming_periodic(nfft, X.dtype)
    Xc = torch.view_as_complex(X.contiguous()) * H.to(X.dtype)[None, :, None]
    fr = torch.fft.irfft(Xc.transpose(1, 2), n=nfft, dim=-1) * w   # (B,M,N)
    L = nfft + h * (M - 1)
    y = torch.zeros(B, L, dtype=X.dtype)
    for m in range(M):
        y[:, m * h:m * h + nfft] += fr[:, m]
    return y / ola_envelope(M, nfft, X.dtype)


def apply_filter(x, H, nfft):
    """utils/blind_bwe_utils.py:6-13."""
    return apply_filter_istft(apply_stft(x, nfft), H, nfft)[:, :x.shape[-1]]


def apply_filter_adjoint(g, H, nfft):
    """Transpose of ``apply_filter`` wrt x: the same
    frame pipeline with the envelope division moved to the input side.
    This is what autograd computes for utils/blind_bwe_utils.py:6-13."""
    B, T = g.shape
    h = nfft // 2
    M = num_frames(T, nfft)
    L = nfft + h * (M - 1)
    w = hamming_periodic(nfft, g.dtype)
    gp = torch.zeros(B, L, dtype=g.dtype)
    gp[:, :T] = g
    gp = gp / ola_envelope(M, nfft, g.dtype)
    fr = gp.unfold(1, nfft, h) * w
    Z = torch.fft.rfft(fr, dim=-1) * H.to(g.dtype)
    fr = torch.fft.irfft(Z, n=nfft, dim=-1) * w
    out = torch.zeros(B, L, dtype=g.dtype)
    for m in range(M):
        out[:, m * h:m * h + nfft] += fr[:, m]
    return out[:, :T]


def apply_filter_grad_H(x, g, nfft):
    """dL/dH for L with dL/d(apply_filter(x,H)) = g:
    sum_{b,m} Re(conj(X) * rfft(w S_m env^-1 g)) * c_k, c_k = 1 at DC/Nyquist
    else 2."""
    B, T = x.shape
    h = nfft // 2
    M = num_frames(T, nfft)
    L = nfft + h * (M - 1)
    w = hamming_periodic(nfft, x.dtype)
    X = torch.view_as_complex(apply_stft(x, nfft))            # (B,F,M)
    gp = torch.zeros(B, L, dtype=g.dtype)
    gp[:, :T] = g
    gp = gp / ola_envelope(M, nfft, g.dtype)
    G = torch.fft.rfft(gp.unfold(1, nfft, h) * w, dim=-1).transpose(1, 2)
    c = torch.full((nfft // 2 + 1,), 2.0, dtype=x.dtype)
    c[0] = 1.0
    c[-1] = 1.0
    return (X.conj() * G).real.sum(dim=(0, 2)) * c / nfft


# --------------------------------------------------------------------------
# a4 / a5 : filter design
# --------------------------------------------------------------------------
def _as_1d(v, dtype):
    if isinstance(v, (list, tuple)):
        v = torch.tensor([float(t) for t in v], dtype=dtype)
    v = torch.as_tensor(v, dtype=dtype)
    return v.reshape(-1)


def design_filter(fc, A, f, G=None):
    """utils/blind_bwe_utils.py:82-119 (and :41-80 when G is given).

    Piecewise dB-per-octave power law.  Segment i covers the bins with
    f >= fc_i; its gain is anchored at the value segment i-1 took at the
    FIRST BIN >= fc_i (not at fc_i itself).  Raises IndexError like the
    reference when some fc_i (i >= 1) exceeds f[-1]."""
    dt = f.dtype
    fc = _as_1d(fc, dt)
    A = _as_1d(A, dt)
    H = torch.zeros_like(f)
    lo = f < fc[0]
    H[lo] = 1
    hi = ~lo
    H[hi] = 10 ** (A[0] * torch.log2(f[hi] / fc[0]) / 20)
    for i in range(1, fc.numel()):
        sel = f >= fc[i]
        idx = torch.nonzero(sel)
        if idx.numel() == 0:
            raise IndexError("index 0 is out of bounds for dimension 0 with size 0")
        anchor = H[idx[0, 0]].clone()
        H[sel] = 10 ** (A[i] * torch.log2(f[sel] / fc[i]) / 20) * anchor
    if G is not None:
        H = H * 10 ** (torch.as_tensor(G, dtype=dt) / 20)
    return H


def design_filter_closed_form(fc, A, f):
    """the same response written without sequential
    overwrites; returns (H, k_first) where k_first[i] is the first bin with
    f >= fc_i.  Used to derive/validate the analytic parameter gradients."""
    dt = f.dtype
    fc = _as_1d(fc, dt)
    A = _as_1d(A, dt)
    K = fc.numel()
    F = f.numel()
    alpha = math.log(10.0) / 20.0
    kf = [int(torch.searchsorted(f, fc[i], right=False)) for i in range(K)]
    # reference overwrites in order i=0..K-1, so the LAST i with f_k >= fc_i wins
    seg = torch.full((F,), -1, dtype=torch.long)
    for i in range(K):
        seg[kf[i]:] = i
    lnH = torch.zeros(F, dtype=dt)
    # anchors: value of segment (owner of bin kf[i] before overwrite) at that bin
    anchor = [torch.zeros((), dtype=dt) for _ in range(K)]
    for i in range(K):
        if i > 0:
            if kf[i] >= F:
                raise IndexError("fc beyond last bin")
            # owner before writing segment i = last j<i with kf[j] <= kf[i]
            j = max([jj for jj in range(i) if kf[jj] <= kf[i]], default=-1)
            if j >= 0:
                anchor[i] = anchor[j] + alpha * A[j] * torch.log2(f[kf[i]] / fc[j])
        sel = seg == i
        lnH[sel] = anchor[i] + alpha * A[i] * torch.log2(f[sel] / fc[i])
    return torch.exp(lnH), kf


def design_filter_vjp(fc, A, f, gH):
    """Analytic (d/dfc, d/dA) of <gH, design_filter(fc,A,f)> following SURVEY
    Appendix A.2 (no gradient flows through the bin indices).  Works for any
    ordering of fc that the sequential overwrite semantics allow."""
    dt = f.dtype
    fc = _as_1d(fc, dt)
    A = _as_1d(A, dt)
    K = fc.numel()
    F = f.numel()
    alpha = math.log(10.0) / 20.0
    H = design_filter(fc, A, f)
    kf = [int(torch.searchsorted(f, fc[i], right=False)) for i in range(K)]
    seg = torch.full((F,), -1, dtype=torch.long)
    for i in range(K):
        seg[kf[i]:] = i
    # chain[i] = list of ancestors (j, bin) contributing to anchor of segment i
    parent = [-1] * K
    for i in range(1, K):
        parent[i] = max([jj for jj in range(i) if kf[jj] <= kf[i]], default=-1)
    gfc = torch.zeros(K, dtype=dt)
    gA = torch.zeros(K, dtype=dt)
    w = gH * H                                     # dL/dlnH_k
    for i in range(K):
        sel = seg == i
        if not bool(sel.any()):
            continue
        wi = w[sel]
        s = wi.sum()
        gA[i] += alpha * (wi * torch.log2(f[sel] / fc[i])).sum()
        gfc[i] += -alpha * A[i] / (fc[i] * math.log(2.0)) * s
        # ancestors contribute through the anchor
        c = i
        while parent[c] >= 0:
            j = parent[c]
            gA[j] += alpha * torch.log2(f[kf[c]] / fc[j]) * s
            gfc[j] += -alpha * A[j] / (fc[j] * math.log(2.0)) * s
            c = j
    return gfc, gA


# --------------------------------------------------------------------------
# frequency weights and losses  (a6, a9, a10, a11)
# --------------------------------------------------------------------------
def freq_weight_vector(kind, F, dtype=torch.float32):
    """The per-bin multiplier the reference applies for each ``freq_weight``
    string (utils/blind_bwe_utils.py:260-293, same table at :162-194 and
    :211-241); freqs = linspace(0,1,F)."""
    fr = torch.linspace(0, 1, F, dtype=dtype)
    if kind == "linear":
        return fr
    if kind == "None":
        return torch.ones(F, dtype=dtype)
    if kind == "log":
        return torch.log2(1 + fr)
    if kind == "sqrt":
        return torch.sqrt(fr)
    if kind == "log2":
        return torch.log2(fr)
    if kind == "log10":
        return torch.log10(fr)
    if kind == "cubic":
        return fr ** 3
    if kind == "quadratic":
        return fr ** 2
    if kind == "logcubic":
        return torch.log2(1 + fr ** 3)
    if kind == "logquadratic":
        return torch.log2(1 + fr ** 2)
    if kind == "squared":
        return fr ** 4
    # the reference silently applies no weighting for unknown strings
    return torch.ones(F, dtype=dtype)


def _mag(X):
    return torch.sqrt(X[..., 0] ** 2 + X[..., 1] ** 2)


def apply_filter_and_norm_STFTmag_fweighted(X, Xref, H, freq_weight="linear"):
    """utils/blind_bwe_utils.py:250-296."""
    w = freq_weight_vector(freq_weight, X.shape[1], X.dtype)[None, :, None]
    a = _mag(X) * H.to(X.dtype)[None, :, None] * w
    b = _mag(Xref) * w
    return torch.linalg.norm(a.reshape(-1) - b.reshape(-1), ord=2)


def apply_filter_and_norm_STFTmag(X, Xref, H):
    """utils/blind_bwe_utils.py:130-141."""
    a = _mag(X) * H.to(X.dtype)[None, :, None]
    return torch.linalg.norm(a.reshape(-1) - _mag(Xref).reshape(-1), ord=2)


def apply_norm_filter(H, H2):
    """utils/blind_bwe_utils.py:143-146."""
    return torch.linalg.norm(H.reshape(-1) - H2.reshape(-1), ord=2)


def apply_norm_STFT_fweighted(y, den_rec, freq_weight="linear", nfft=1024):
    """utils/blind_bwe_utils.py:148-197 (weights act on re and im parts)."""
    X = apply_stft(den_rec, nfft)
    Xr = apply_stft(y, nfft)
    w = freq_weight_vector(freq_weight, X.shape[1], X.dtype)[None, :, None, None]
    return torch.linalg.norm((X * w).reshape(-1) - (Xr * w).reshape(-1), ord=2)


def apply_norm_STFTmag_fweighted(y, den_rec, freq_weight="linear", nfft=1024,
                                 logmag=False):
    """utils/blind_bwe_utils.py:198-248."""
    X = _mag(apply_stft(den_rec, nfft))
    Xr = _mag(apply_stft(y, nfft))
    w = freq_weight_vector(freq_weight, X.shape[1], X.dtype)[None, :, None]
    X = X * w
    Xr = Xr * w
    if logmag:
        return torch.linalg.norm(torch.log10(X.reshape(-1) + 1e-8)
                                 - torch.log10(Xr.reshape(-1) + 1e-8), ord=2)
    return torch.linalg.norm(X.reshape(-1) - Xr.reshape(-1), ord=2)


# --------------------------------------------------------------------------
# fit statistics
# --------------------------------------------------------------------------
def stft_mag_stats(x, y, nfft):
    """a_k = sum |X|^2, b_k = sum |X||Y|, c_k = sum |Y|^2 over batch and frames,
    X = STFT(x), Y = STFT(y).  The weighted STFT-magnitude loss of
    utils/blind_bwe_utils.py:250-296 is sqrt(sum_k w_k^2 (H_k^2 a_k - 2 H_k b_k
    + c_k))."""
    X = _mag(apply_stft(x, nfft))
    Y = _mag(apply_stft(y, nfft))
    return (X * X).sum(dim=(0, 2)), (X * Y).sum(dim=(0, 2)), (Y * Y).sum(dim=(0, 2))


def norm_from_stats(a, b, c, H, w):
    S = (w * w * (H * H * a - 2 * H * b + c)).sum()
    return torch.sqrt(torch.clamp(S, min=0))


# --------------------------------------------------------------------------
# reconstruction guidance, operator part (a8)
# --------------------------------------------------------------------------
def rec_guidance_operator(x_hat, y, H, nfft):
    """norm_b = ||y_b - A(x_hat_b)||_2 and the cotangent the reference's
    autograd feeds into the denoiser backward:
    d(sum_b norm_b)/d x_hat = A^T((A x_hat - y)/norm_b)
    (testing/blind_bwe_sampler.py:89,117,120)."""
    r = apply_filter(x_hat, H, nfft) - y
    n = torch.linalg.norm(r, dim=1)
    g = apply_filter_adjoint(r / n[:, None], H, nfft)
    return n, g
