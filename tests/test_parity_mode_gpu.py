"""GPU: the drop-in functions driven exactly the way the UNCHANGED reference sampler drives them
(testing/blind_bwe_sampler.py:522-595: design_filter -> weighted STFT-magnitude norm ->
autograd.grad(create_graph=True) -> step -> in-place clamps), plus randomised cases
(SURVEY section 4 'hypothesis' row) against the oracle."""
import numpy as np
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bu():
    from babe_b200 import build
    build.build()
    from babe_b200 import blind_bwe_utils
    return blind_bwe_utils


def test_reference_fit_loop_call_pattern(bu, golden):
    g = golden("fit_sampler.npz")
    nfft, sr = int(g["nfft"]), int(g["sr"])
    xden, y = torch.from_numpy(g["fit_xden"]).cuda(), torch.from_numpy(g["y"]).cuda()
    filter_params = torch.from_numpy(g["fit_p0"]).cuda()
    freqs = torch.fft.rfftfreq(nfft, d=1 / sr).cuda()
    mu = torch.Tensor([1000, 10]).cuda()
    fcmin, fcmax, Amin = 20, sr // 2, -50
    # --- lines :556-:583 of the reference, verbatim call pattern ---
    Xden = bu.apply_stft(xden, nfft)
    Y = bu.apply_stft(y, nfft)
    for i in range(5):
        filter_params.requires_grad = True
        H = bu.design_filter(filter_params[0], filter_params[1], freqs)
        norm = bu.apply_filter_and_norm_STFTmag_fweighted(Xden, Y, H, "sqrt")
        grad = torch.autograd.grad(norm, filter_params, create_graph=True)
        filter_params = filter_params - mu.unsqueeze(1) * grad[0]
        filter_params.detach_()
        filter_params[0, 0] = torch.clamp(filter_params[0, 0], min=fcmin, max=fcmax)
        for k in range(1, len(filter_params[0])):
            filter_params[0, k] = torch.clamp(filter_params[0, k], min=filter_params[0, k - 1] + 1, max=fcmax)
        filter_params[1, 0] = torch.clamp(filter_params[1, 0], min=Amin, max=-1)
        for k in range(1, len(filter_params[0])):
            filter_params[1, k] = torch.clamp(filter_params[1, k], min=Amin, max=filter_params[1, k - 1])
        if i == 0:
            assert rel_l2(filter_params.cpu(), g["fit_p_1"]) < 1e-5
    assert rel_l2(filter_params.cpu(), g["fit_p_5"]) < 1e-5


@pytest.mark.parametrize("seed", range(6))
def test_randomised_filters_and_shapes(bu, seed):
    from oracle import stft_filter as sf
    rng = np.random.default_rng(seed)
    nfft = int(rng.choice([1024, 4096]))
    sr = int(rng.choice([22050, 44100]))
    B, T = int(rng.integers(1, 5)), int(rng.integers(1, 6 * nfft))
    K = int(rng.integers(1, 8))
    f = torch.fft.rfftfreq(nfft, d=1 / sr)
    fc = np.sort(rng.uniform(20, sr / 2, K)).astype(np.float32)
    if seed % 3 == 0 and K >= 2:
        fc[1] = fc[0] + 0.25 * float(f[1])            # two breakpoints inside one bin
    if seed % 3 == 1:
        fc[-1] = float(f[-1])                         # breakpoint exactly at Nyquist
    A = -np.sort(rng.uniform(1, 50, K)).astype(np.float32)
    fc_t, A_t = torch.from_numpy(fc), torch.from_numpy(A)
    H = sf.design_filter(fc_t, A_t, f)
    Hc = bu.design_filter(fc_t.cuda(), A_t.cuda(), f.cuda())
    assert rel_l2(Hc.cpu(), H) < 1e-5
    x = torch.from_numpy(rng.standard_normal((B, T)).astype(np.float32)) * 0.1
    assert rel_l2(bu.apply_filter(x.cuda(), Hc, nfft).cpu(), sf.apply_filter(x, H, nfft)) < 1e-5
    # VJP through design_filter on a random cotangent
    cot = torch.from_numpy(rng.standard_normal(f.shape).astype(np.float32))
    gfc, gA = sf.design_filter_vjp(fc_t.double(), A_t.double(), f.double(), cot.double())
    fcg, Ag = fc_t.cuda().requires_grad_(True), A_t.cuda().requires_grad_(True)
    g1, g2 = torch.autograd.grad((bu.design_filter(fcg, Ag, f.cuda()) * cot.cuda()).sum(), (fcg, Ag))
    assert rel_l2(g1.cpu(), gfc) < 1e-4 and rel_l2(g2.cpu(), gA) < 1e-4
