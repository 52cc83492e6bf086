"""CPU: the Python plumbing of BlindSamplerFused.predict_blind_bwe (schedule, stochastic step, fit
call, fused rec-guidance Function, Heun correction, in-place filter update) with the three C-ABI
operators it uses replaced by their oracle restatements, against the trajectory of the UNMODIFIED
reference sampler (tests/golden/fit_sampler.npz: toy denoiser, 4 steps, 20 fit iterations, seed 42).
The kernels themselves are checked on the GPU (tests/test_sampler_gpu.py)."""
import pytest
import torch

from conftest import rel_l2
from toy_model import ToyDenoiser


@pytest.fixture
def oracle_ops(monkeypatch):
    from babe_b200 import blind_bwe_utils as bu, ops
    from oracle import filter_fit as ofit, stft_filter as osf

    def stft_stats(x, y, nfft, mode=0):
        assert mode == 0
        return torch.stack(osf.stft_mag_stats(x.double(), y.double(), nfft))

    def fit_params(abc, w, freqs, params, cfg, return_iters=False):
        sr = int(round(float(freqs[1]) * 2 * (freqs.numel() - 1)))
        oc = ofit.FitConfig(nfft=2 * (freqs.numel() - 1), sample_rate=sr, fcmin=cfg.fcmin, fcmax=cfg.fcmax,
                            Amin=cfg.Amin, Amax=cfg.Amax, max_iter=cfg.max_iter, tol=(cfg.tol_fc, cfg.tol_A),
                            mu=(cfg.mu_fc, cfg.mu_A), clamp_fc=bool(cfg.clamp_fc), clamp_A=bool(cfg.clamp_A),
                            only_negative_A=bool(cfg.only_negative_A))
        p, it = ofit.fit_params_from_stats(abc[0].float(), abc[1].float(), abc[2].float(), params, oc)
        params.copy_(p)                                               # in place, like the kernel
        return (params, torch.tensor([it])) if return_iters else params

    def apply_filter(x, nfft, H=None, freqs=None, fc=None, A=None, adjoint=False, sub=None, row_scale=None,
                     row_sumsq=None, out=None):
        if H is None:
            H = osf.design_filter(fc, A, freqs)
        y = osf.apply_filter_adjoint(x, H, nfft) if adjoint else osf.apply_filter(x, H, nfft)
        if sub is not None:
            y = y - sub
        if row_sumsq is not None:
            row_sumsq += (y.double() ** 2).sum(1)
        if row_scale is not None:
            y = y * row_scale[:, None]
        return y

    monkeypatch.setattr(ops, "stft_stats", stft_stats)
    monkeypatch.setattr(ops, "fit_params", fit_params)
    monkeypatch.setattr(ops, "apply_filter", apply_filter)
    monkeypatch.setattr(bu, "freq_weight_vector", lambda kind, F, device: osf.freq_weight_vector(kind, F))


def test_blind_sampler_plumbing_matches_reference_golden(oracle_ops, golden):
    from babe_b200 import edm, sampler
    g = golden("fit_sampler.npz")
    y = torch.from_numpy(g["y"])
    args = sampler.make_args(sample_rate=int(g["sr"]), audio_len=y.shape[1], T=4, NFFT=int(g["nfft"]), max_iter=20)
    s = sampler.BlindSamplerFused(ToyDenoiser(), edm.EDM(args), args, rid=False)
    torch.manual_seed(42)
    x, p = s.predict_blind_bwe(y.clone())
    # same tolerance as the GPU test: chained fp32 fits amplify rounding
    assert rel_l2(p, g["sampler_params"]) < 1e-2
    assert rel_l2(x, g["sampler_x"]) < 1e-2
    # rid=True returns the reference's 5-tuple
    torch.manual_seed(42)
    out = s.predict_blind_bwe(y.clone(), rid=True, max_steps=2)
    assert len(out) == 5 and out[2].shape == (4, *y.shape)
