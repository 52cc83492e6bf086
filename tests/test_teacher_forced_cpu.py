"""CPU: the oracle's filter fit (oracle/filter_fit.py) started from each of the reference's 100 iterates
(tests/golden/teacher_forced.npz, produced by the unmodified BlindSampler.fit_params) reproduces the reference's next
iterate -- pins the oracle's update rule, clamps and statistics decomposition iteration by iteration."""
import numpy as np
import pytest
import torch

from conftest import rel_l2


@pytest.mark.parametrize("case,inputs,xk,yk", [("n1024", "fit_sampler.npz", "fit_xden", "y"),
                                                ("k7", "fit_sampler.npz", "fit_xden", "y"),
                                                ("n4096", "operator_n4096.npz", "x", "yobs")])
def test_oracle_fit_iteration_from_every_reference_iterate(golden, case, inputs, xk, yk):
    from oracle import filter_fit as ofit, stft_filter as osf
    tf, g = golden("teacher_forced.npz"), golden(inputs)
    its = tf[f"fit_iters_{case}"]
    nfft, sr = int(g["nfft"]), int(g["sr"])
    cfg = ofit.FitConfig(nfft=nfft, sample_rate=sr, max_iter=1)
    a, b, c = osf.stft_mag_stats(torch.from_numpy(g[xk]).double(), torch.from_numpy(g[yk]).double(), nfft)
    worst = 0.0
    for i in range(0, its.shape[0] - 1):
        p, _ = ofit.fit_params_from_stats(a, b, c, torch.from_numpy(its[i]).double(), cfg, dtype=torch.float64)
        worst = max(worst, rel_l2(p[0], its[i + 1][0]), rel_l2(p[1], its[i + 1][1]))
    assert worst < 1e-5, (case, worst)
    assert int(tf[f"fit_stop_{case}"]) >= 1
