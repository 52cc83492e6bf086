"""Oracle: per-step filter re-estimation of the blind sampler (CPU, torch).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates
``BlindSampler.fit_params`` / ``optimizer_func``
(testing/blind_bwe_sampler.py:522-595 of eloimoliner/BABE): projected
gradient descent on the breakpoints (fc, A) of the parametric lowpass so that
``H * |STFT(x_den)|`` matches ``|STFT(y)|`` under a frequency weighting.
"""
import math
from dataclasses import dataclass, field

import torch

from . import stft_filter as sf


@dataclass
class FitConfig:
    """Values read from ``args.tester.blind_bwe`` / ``args.exp`` by
    BlindSampler.__init__ (testing/blind_bwe_sampler.py:31-41) with the
    defaults of conf/tester/blind_bwe.yaml:129-153."""
    nfft: int = 4096
    sample_rate: int = 22050
    fcmin: float = 20.0
    fcmax: object = "nyquist"
    Amin: float = -50.0
    Amax: float = 30.0
    max_iter: int = 100
    tol: tuple = (5e-3, 5e-3)
    mu: tuple = (1000.0, 10.0)
    clamp_fc: bool = True
    clamp_A: bool = True
    only_negative_A: bool = True
    freq_weighting_filter: str = "sqrt"

    def fcmax_value(self):
        # testing/blind_bwe_sampler.py:35-38: integer division
        return self.sample_rate // 2 if self.fcmax == "nyquist" else self.fcmax


def rfft_freqs(nfft, sample_rate, dtype=torch.float32):
    """testing/blind_bwe_sampler.py:629."""
    return torch.fft.rfftfreq(nfft, d=1 / sample_rate).to(dtype)


def project_params(p, cfg):
    """The sequential clamps of testing/blind_bwe_sampler.py:576-583, in place
    on p (2,K)."""
    K = p.shape[1]
    fcmax = cfg.fcmax_value()
    if cfg.clamp_fc:
        p[0, 0] = torch.clamp(p[0, 0], min=cfg.fcmin, max=fcmax)
        for k in range(1, K):
            p[0, k] = torch.clamp(p[0, k], min=p[0, k - 1] + 1, max=fcmax)
    if cfg.clamp_A:
        p[1, 0] = torch.clamp(p[1, 0], min=cfg.Amin,
                              max=-1 if cfg.only_negative_A else cfg.Amax)
        for k in range(1, K):
            # torch.clamp(min, max) with min > max returns max
            hi = p[1, k - 1] if cfg.only_negative_A else torch.tensor(cfg.Amax, dtype=p.dtype)
            p[1, k] = torch.minimum(torch.maximum(p[1, k], torch.tensor(cfg.Amin, dtype=p.dtype)), hi)
    return p


def loss_and_grad_from_stats(a, b, c, params, f, w):
    """norm = sqrt(S), S = sum_k w_k^2 (H_k^2 a_k - 2 H_k b_k + c_k), and its
    analytic gradient wrt (fc, A)."""
    H = sf.design_filter(params[0], params[1], f)
    w2 = w * w
    S = (w2 * (H * H * a - 2 * H * b + c)).sum()
    norm = torch.sqrt(S)
    gH = w2 * (H * a - b) / norm                     # d norm / d H_k
    gfc, gA = sf.design_filter_vjp(params[0], params[1], f, gH)
    return norm, torch.stack((gfc, gA))


def fit_params_from_stats(a, b, c, params, cfg, dtype=None):
    """testing/blind_bwe_sampler.py:562-590 with the loss evaluated through
    the three F-vectors (a,b,c).  Returns (params, iterations_run)."""
    dtype = dtype or params.dtype
    p = params.detach().clone().to(dtype)
    F = a.numel()
    f = rfft_freqs(cfg.nfft, cfg.sample_rate, torch.float32).to(dtype)
    w = sf.freq_weight_vector(cfg.freq_weighting_filter, F, torch.float32).to(dtype)
    a, b, c = a.to(dtype), b.to(dtype), c.to(dtype)
    mu = torch.tensor(cfg.mu, dtype=dtype)[:, None]
    prev = None
    it = 0
    for i in range(cfg.max_iter):
        _, g = loss_and_grad_from_stats(a, b, c, p, f, w)
        p = p - mu * g
        p = project_params(p, cfg)
        it = i + 1
        if i > 0:
            if (p[0] - prev[0]).abs().mean() < cfg.tol[0] and \
               (p[1] - prev[1]).abs().mean() < cfg.tol[1]:
                break
        prev = p.clone()
    return p, it


def fit_params_literal(x_den, y, params, cfg):
    """Closest restatement of fit_params: autograd through design_filter and
    the direct (not a,b,c-collapsed) loss, as testing/blind_bwe_sampler.py
    does.  Used to pin ``fit_params_from_stats``."""
    p = params.detach().clone()
    f = rfft_freqs(cfg.nfft, cfg.sample_rate, p.dtype)
    Xd = sf.apply_stft(x_den, cfg.nfft)
    Y = sf.apply_stft(y, cfg.nfft)
    mu = torch.tensor(cfg.mu, dtype=p.dtype)[:, None]
    prev = None
    it = 0
    for i in range(cfg.max_iter):
        p.requires_grad_(True)
        H = sf.design_filter(p[0], p[1], f)
        norm = sf.apply_filter_and_norm_STFTmag_fweighted(Xd, Y, H, cfg.freq_weighting_filter)
        (g,) = torch.autograd.grad(norm, p)
        p = (p - mu * g).detach()
        p = project_params(p, cfg)
        it = i + 1
        if i > 0:
            if (p[0] - prev[0]).abs().mean() < cfg.tol[0] and \
               (p[1] - prev[1]).abs().mean() < cfg.tol[1]:
                break
        prev = p.clone()
    return p, it


def fit_params(x_den, y, params, cfg):
    """STFT statistics + device-style fit loop, the decomposition the CUDA
    path uses."""
    a, b, c = sf.stft_mag_stats(x_den, y, cfg.nfft)
    return fit_params_from_stats(a, b, c, params, cfg)
