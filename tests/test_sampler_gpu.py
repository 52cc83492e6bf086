"""GPU: the fused blind sampler against the golden trajectory the UNMODIFIED
reference sampler produced (tests/golden/make_golden.py: toy denoiser, 4 steps,
20 fit iterations, torch.manual_seed(42), host noise) and against the oracle
restatement of the loop."""
import numpy as np
import pytest
import torch

from conftest import rel_l2
from toy_model import ToyDenoiser

pytestmark = pytest.mark.gpu


def _setup(golden, max_iter=20):
    from babe_b200 import build
    build.build()
    from babe_b200 import edm, sampler
    g = golden("fit_sampler.npz")
    y = torch.from_numpy(g["y"]).cuda()
    args = sampler.make_args(sample_rate=int(g["sr"]), audio_len=y.shape[1], T=4, NFFT=int(g["nfft"]),
                             max_iter=max_iter)
    model = ToyDenoiser().cuda()
    s = sampler.BlindSamplerFused(model, edm.EDM(args), args, rid=False)
    return g, y, args, model, s


def test_fused_sampler_matches_reference_golden(golden):
    g, y, args, model, s = _setup(golden)
    torch.manual_seed(42)
    x, p = s.predict_blind_bwe(y.clone())
    # chained fp32 gradient-descent fits amplify rounding differences (see
    # tests/test_oracle_golden.py::test_fit_params); the oracle restatement with
    # the same (a,b,c) decomposition deviates from the golden by the same amount
    assert rel_l2(p.cpu(), g["sampler_params"]) < 1e-2
    assert rel_l2(x.cpu(), g["sampler_x"]) < 1e-2


def test_fused_sampler_matches_oracle_loop(golden):
    """Same seeds, same decomposition (statistics + analytic fit) on CPU."""
    from oracle import blind_sampler as obs, filter_fit as ofit
    g, y, args, model, s = _setup(golden, max_iter=3)
    torch.manual_seed(7)
    x, p = s.predict_blind_bwe(y.clone())
    cfg = obs.SamplerConfig(T=4, audio_len=y.shape[1])
    cfg.fit = ofit.FitConfig(nfft=int(g["nfft"]), sample_rate=int(g["sr"]), max_iter=3)
    cpu_model = ToyDenoiser()
    torch.manual_seed(7)
    xo, po = obs.predict_blind_bwe(cfg, cpu_model, cpu_model.CQTransform.apply_hpf_DC, y.cpu())
    assert rel_l2(p.cpu(), po) < 1e-4
    assert rel_l2(x.cpu(), xo) < 1e-4


def test_rid_outputs_and_device_noise(golden):
    g, y, args, model, s = _setup(golden, max_iter=2)
    torch.manual_seed(0)
    out = s.predict_blind_bwe(y.clone(), rid=True)
    assert len(out) == 5
    x, p, den, t, filt = out
    assert den.shape == (4, *y.shape) and filt.shape == (4, 2, 5) and t.shape == (5,)
    assert torch.isfinite(x).all()
    # every babe_b200 kernel is bitwise reproducible (fixed-order reductions, no
    # atomics); cuDNN's conv backward in the toy denoiser needs its deterministic mode
    torch.backends.cudnn.deterministic = True
    s.device_noise = True
    s.generator = torch.Generator(device="cuda").manual_seed(3)
    x1, _ = s.predict_blind_bwe(y.clone())
    s.generator = torch.Generator(device="cuda").manual_seed(3)
    x2, _ = s.predict_blind_bwe(y.clone())
    assert torch.equal(x1, x2)


def test_cuda_graph_replay_matches_eager(golden):
    """the denoiser's forward/backward replayed from captured CUDA graphs."""
    from babe_b200 import profiling
    g, y, args, model, s = _setup(golden, max_iter=3)
    torch.backends.cudnn.deterministic = True
    torch.manual_seed(5)
    x0, p0 = s.predict_blind_bwe(y.clone(), max_steps=3)
    s.cuda_graph = True
    torch.manual_seed(5)
    n0 = profiling.launches()
    x1, p1 = s.predict_blind_bwe(y.clone(), max_steps=3)
    assert s._graphed is not None, "capture failed"
    assert profiling.launches() > n0
    assert rel_l2(x1, x0) < 1e-6 and rel_l2(p1, p0) < 1e-6


def test_whole_step_graph_matches_eager(golden):
    """ONE captured CUDA graph per sampler step (move, denoiser fwd + bwd, statistics, fit loop, fused
    guidance, Heun correction, state update) replayed with device-resident schedule scalars."""
    g, y, args, model, s = _setup(golden, max_iter=3)
    torch.backends.cudnn.deterministic = True
    torch.manual_seed(5)
    x0, p0, den0, t0, f0 = s.predict_blind_bwe(y.clone(), rid=True)
    s.step_graph = True
    torch.manual_seed(5)
    x1, p1, den1, t1, f1 = s.predict_blind_bwe(y.clone(), rid=True)
    assert s._step_graphs, "capture failed"
    assert rel_l2(x1, x0) < 1e-6 and rel_l2(p1, p0) < 1e-6
    assert rel_l2(den1, den0) < 1e-6 and rel_l2(f1, f0) < 1e-6
    torch.manual_seed(5)                                   # replay again from the cached graphs
    x2, p2 = s.predict_blind_bwe(y.clone())
    assert rel_l2(x2, x0) < 1e-6 and rel_l2(p2, p0) < 1e-6


def test_compute_sweep_matches_oracle(golden):
    """a14 (testing/blind_bwe_sampler.py:598-616): loss and (d/dfc, d/dA) on the 15x12 grid."""
    from oracle import filter_fit as ofit, stft_filter as osf
    from babe_b200 import sampler
    g, y, args, model, s = _setup(golden, max_iter=2)
    dev = y.device
    s.freqs = torch.fft.rfftfreq(args.tester.blind_bwe.NFFT, d=1 / args.exp.sample_rate).to(dev)
    s._fit = sampler.FilterFit.from_args(args, dev)
    s.fc_s = torch.logspace(2.5, 4, 15).to(dev)
    s.A_s = torch.linspace(-80, -5, 12).to(dev)
    xden = torch.from_numpy(g["fit_xden"]).to(dev)
    norms, grads = s.compute_sweep(xden, y)
    assert norms.shape == (15, 12) and grads.shape == (15, 12, 2)
    a, b, c = osf.stft_mag_stats(xden.cpu().double(), y.cpu().double(), int(g["nfft"]))
    f = ofit.rfft_freqs(int(g["nfft"]), int(g["sr"]), torch.float64)
    w = osf.freq_weight_vector("sqrt", f.numel()).double()
    for i, j in ((0, 0), (7, 5), (14, 11)):
        p = torch.stack((s.fc_s[i].reshape(1), s.A_s[j].reshape(1))).cpu().double()
        nrm, gr = ofit.loss_and_grad_from_stats(a, b, c, p, f, w)
        assert abs(float(norms[i, j]) - float(nrm)) < 1e-4 * float(nrm)
        assert rel_l2(grads[i, j], gr.reshape(-1)) < 1e-3


@pytest.mark.parametrize("variant", ["data_consistency", "snr", "smoothl1", "cosine", "stft", "stft_mag",
                                     "stft_logmag", "sweep_rid"])
def test_optional_branches_run(golden, variant):
    """Non-default branches of get_rec_grads / fit_params / the sampling loop
    (testing/blind_bwe_sampler.py:80-115, 542-548, 704-709): executed on the CUDA operators."""
    g, y, args, model, s = _setup(golden, max_iter=2)
    ps = args.tester.posterior_sampling
    y_in = y.clone()
    kw = {}
    if variant == "data_consistency":
        ps.data_consistency = True
        s.data_consistency = True
    elif variant == "snr":
        ps.SNR_observations = 30
    elif variant in ("smoothl1", "cosine"):
        ps.norm = variant
    elif variant.startswith("stft"):
        # a frame that is ALL zero padding has |X| = 0 and sqrt'(0) = inf -> NaN gradients, in the
        # reference exactly as here (utils/blind_bwe_utils.py:206-207); avoid T % hop == 0
        y = y[:, :4000].contiguous()
        y_in = y.clone()
        args.exp.audio_len = 4000
        ps.stft_distance.use = True
        ps.stft_distance.nfft = 1024
        ps.freq_weighting = "sqrt"
        ps.stft_distance.mag = variant != "stft"
        ps.stft_distance.logmag = variant == "stft_logmag"
    elif variant == "sweep_rid":
        kw = dict(rid=True, compute_sweep=True)
    torch.manual_seed(1)
    out = s.predict_blind_bwe(y_in, max_steps=2, **kw)
    x = out[0]
    assert x.shape == y.shape and torch.isfinite(x).all()
    if variant == "snr":
        assert not torch.equal(y_in, y)                  # the reference adds noise to y IN PLACE (:86, :548)
    if variant == "sweep_rid":
        assert len(out) == 7 and out[5].shape == (4, 15, 12) and out[6].shape == (4, 15, 12, 2)


def test_non_blind_and_unconditional(golden):
    g, y, args, model, s = _setup(golden, max_iter=2)
    filt = torch.tensor([[1000.0], [-20.0]]).cuda()
    x = s.predict_bwe(y.clone(), filt, "fc_A")
    assert x.shape == y.shape and torch.isfinite(x).all()
    from babe_b200 import bandwidth_extension as bwe
    taps = bwe.get_FIR_lowpass(500, 1000, 1, 22050).cuda()
    xf = s.predict_bwe(bwe.apply_low_pass_firwin(y, taps), taps, "firwin")
    assert xf.shape == y.shape and torch.isfinite(xf).all()
    with pytest.raises(NotImplementedError):
        s.predict_bwe(y.clone(), filt, "cheby1")
    xu = s.predict_unconditional(tuple(y.shape), y.device)
    assert xu.shape == y.shape and torch.isfinite(xu).all()


def test_predict_bwe_AR_and_bwe_match_reference(golden):
    """Non-blind and autoregressive sampling (testing/blind_bwe_sampler.py:259-364, 406-497) on the CUDA
    operators against trajectories of the UNMODIFIED reference sampler (tests/golden/make_golden_ar.py)."""
    from babe_b200 import build
    build.build()
    from babe_b200 import edm, sampler
    g = golden("sampler_ar.npz")
    c = lambda k: torch.from_numpy(g[k]).cuda()
    args = sampler.make_args(sample_rate=int(g["sr"]), audio_len=g["x"].shape[1], T=4, NFFT=int(g["nfft"]),
                             max_iter=20)
    model = ToyDenoiser().cuda()
    s = sampler.BlindSamplerFused(model, edm.EDM(args), args, rid=False)
    torch.manual_seed(77)
    x = s.predict_bwe_AR(c("ylpf"), c("y_masked"), c("filt"), "fc_A", mask=c("mask"))
    assert rel_l2(x.cpu(), g["x_ar"]) < 1e-3
    s2 = sampler.BlindSamplerFused(model, edm.EDM(args), args, rid=False)
    torch.manual_seed(78)
    xb = s2.predict_bwe(c("ylpf"), c("filt"), "fc_A")
    assert rel_l2(xb.cpu(), g["x_bwe"]) < 1e-3
