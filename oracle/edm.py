"""Oracle: Karras EDM schedule and preconditioning (CPU, torch).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates the parts of
``diff_params/edm.py`` (eloimoliner/BABE) that the blind sampler evaluates.
"""
import math
from dataclasses import dataclass

import torch


@dataclass
class EDMConfig:
    """conf/tester/blind_bwe.yaml:56-68 (the tester block overrides the
    training values, testing/blind_bwe_sampler.py:50-60)."""
    sigma_data: float = 0.063
    sigma_min: float = 1e-4
    sigma_max: float = 1.0
    ro: float = 8.0
    Schurn: float = 20.0
    Snoise: float = 1.0
    Stmin: float = 0.0
    Stmax: float = 50.0


def create_schedule_from_initial_t(cfg, initial_t, nb_steps):
    """diff_params/edm.py:66-75.  nb_steps+1 points, abscissa i/(nb_steps-1),
    last entry overwritten with 0."""
    i = torch.arange(0, nb_steps + 1)
    t = (initial_t ** (1 / cfg.ro) + i / (nb_steps - 1)
         * (cfg.sigma_min ** (1 / cfg.ro) - initial_t ** (1 / cfg.ro))) ** cfg.ro
    t[-1] = 0
    return t


def create_schedule(cfg, nb_steps):
    """diff_params/edm.py:55-64."""
    return create_schedule_from_initial_t(cfg, cfg.sigma_max, nb_steps)


def get_gamma(cfg, t):
    """diff_params/edm.py:38-53."""
    N = t.shape[0]
    gamma = torch.zeros(t.shape)
    sel = torch.logical_and(t > cfg.Stmin, t < cfg.Stmax)
    gamma[sel] = gamma[sel] + torch.min(torch.Tensor([cfg.Schurn / N, 2 ** (1 / 2) - 1]))
    return gamma


def cskip(cfg, sigma):
    return cfg.sigma_data ** 2 * (sigma ** 2 + cfg.sigma_data ** 2) ** -1


def cout(cfg, sigma):
    return sigma * cfg.sigma_data * (cfg.sigma_data ** 2 + sigma ** 2) ** (-0.5)


def cin(cfg, sigma):
    return (cfg.sigma_data ** 2 + sigma ** 2) ** (-0.5)


def cnoise(cfg, sigma):
    return (1 / 4) * torch.log(sigma)


def denoiser(cfg, xn, net, sigma):
    """diff_params/edm.py:144-159: cskip*x + cout*net(cin*x, cnoise)."""
    if sigma.dim() == 1:
        sigma = sigma.unsqueeze(-1)
    return cskip(cfg, sigma) * xn + cout(cfg, sigma) * net(cin(cfg, sigma) * xn, cnoise(cfg, sigma))
