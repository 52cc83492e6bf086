"""Oracle: the blind posterior-sampling loop (CPU, torch).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates
``BlindSampler.predict_blind_bwe`` and the helpers it calls
(testing/blind_bwe_sampler.py:75-157, 509-595, 619-769 of eloimoliner/BABE)
on top of the oracle operator, for the default branch choices of
conf/tester/blind_bwe.yaml (norm 2, no STFT distance, no data consistency, no
observation noise, rid False).
"""
from dataclasses import dataclass, field

import torch

from . import edm as oedm
from . import filter_fit as ofit
from . import stft_filter as sf


@dataclass
class SamplerConfig:
    """conf/tester/blind_bwe.yaml:21-53,129-153; conf/exp/maestro22k_8s.yaml:51-52."""
    T: int = 35
    order: int = 2
    xi: float = 0.2
    start_sigma: float = 0.2
    audio_len: int = 184184
    filter_out_cqt_DC_Nyq: bool = True
    fc_init: tuple = (280.0, 285.0, 290.0, 295.0, 300.0)
    A_init: tuple = (-15.0, -17.0, -20.0, -25.0, -30.0)
    edm: oedm.EDMConfig = field(default_factory=oedm.EDMConfig)
    fit: ofit.FitConfig = field(default_factory=ofit.FitConfig)


def move_timestep(x, t, gamma, Snoise=1.0, randn=torch.randn):
    """testing/blind_bwe_sampler.py:509-516 -- always draws, even if gamma==0."""
    t_hat = t + gamma * t
    eps = randn(x.shape) * Snoise
    return x + ((t_hat ** 2 - t ** 2) ** (1 / 2)) * eps, t_hat


def denoised_estimate(cfg, net, hpf, x, t_i):
    """testing/blind_bwe_sampler.py:152-157."""
    x_hat = oedm.denoiser(cfg.edm, x, net, t_i.unsqueeze(-1))
    if cfg.filter_out_cqt_DC_Nyq:
        x_hat = hpf(x_hat)
    return x_hat


def rec_grads(cfg, x_den, y, x, t_i, H):
    """testing/blind_bwe_sampler.py:75-135 (default branch)."""
    den_rec = sf.apply_filter(x_den, H, cfg.fit.nfft)
    norm = torch.linalg.norm(y - den_rec, dim=1, ord=2)
    (g,) = torch.autograd.grad(norm.sum(), x)
    normguide = torch.linalg.norm(g) / cfg.audio_len ** 0.5
    s = cfg.xi / (normguide + 1e-6)
    return s * g / t_i


def one_evaluation(cfg, net, hpf, x_in, t_in, y, filter_params, f, fit_fn):
    """One {denoise, fit, guidance} evaluation: lines :689-:701 (and again
    :735-:745 for the Heun correction)."""
    x_in = x_in.detach().requires_grad_(True)
    x_den = denoised_estimate(cfg, net, hpf, x_in, t_in)
    x_den_2 = x_den.clone().detach()
    filter_params, _ = fit_fn(x_den_2, y, filter_params, cfg.fit)
    H = sf.design_filter(filter_params[0], filter_params[1], f)
    g = rec_grads(cfg, x_den, y, x_in, t_in, H)
    x_in = x_in.detach()
    score = (x_den_2 - x_in) / t_in ** 2 - g
    return score, filter_params, x_in


def predict_blind_bwe(cfg, net, hpf, y, steps=None, fit_fn=None, randn=torch.randn,
                      trace=None):
    """testing/blind_bwe_sampler.py:619-769.  ``steps`` bounds the number of
    loop iterations actually run (the schedule is always built for cfg.T);
    ``trace`` (a list) receives per-step (x, filter_params) when given."""
    fit_fn = fit_fn or ofit.fit_params
    f = ofit.rfft_freqs(cfg.fit.nfft, cfg.fit.sample_rate)
    filter_params = torch.Tensor([list(cfg.fc_init), list(cfg.A_init)])
    t = oedm.create_schedule_from_initial_t(cfg.edm, cfg.start_sigma, cfg.T)
    x = y + randn(y.shape) * t[0]
    gamma = oedm.get_gamma(cfg.edm, t)
    n = cfg.T if steps is None else min(steps, cfg.T)
    for i in range(n):
        x_hat, t_hat = move_timestep(x, t[i], gamma[i], randn=randn)
        score, filter_params, x_hat = one_evaluation(cfg, net, hpf, x_hat, t_hat, y,
                                                     filter_params, f, fit_fn)
        d = -t_hat * score
        h = t[i + 1] - t_hat
        if t[i + 1] != 0 and cfg.order == 2:
            t_prime = t[i + 1]
            x_prime = x_hat + h * d
            score, filter_params, _ = one_evaluation(cfg, net, hpf, x_prime, t_prime, y,
                                                     filter_params, f, fit_fn)
            d_prime = -t_prime * score
            x = x_hat + h * ((1 / 2) * d + (1 / 2) * d_prime)
        else:
            x = x_hat + h * d
        if trace is not None:
            trace.append((x.detach().clone(), filter_params.detach().clone()))
    return x.detach(), filter_params.detach()
