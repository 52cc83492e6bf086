"""Karras EDM schedule and preconditioning used by the blind sampler.

Host-side scalar mathematics mirroring ``diff_params/edm.py`` (EDM class) of
eloimoliner/BABE so that ``args.diff_params.callable`` can point here when the
reference tree is not importable (e.g. on the GPU box).  Same attribute and
method names, including the training-side members (noise-level sampling, preconditioned targets,
``loss_fn``, which calls ``CQT_nsgt.apply_hpf_DC`` at diff_params/edm.py:197).
"""
import numpy as np
import torch


class EDM:
    def __init__(self, args):
        """diff_params/edm.py:11-35."""
        self.args = args
        d = args.diff_params
        self.sigma_min = d.sigma_min
        self.sigma_max = d.sigma_max
        self.ro = d.ro
        self.sigma_data = d.sigma_data
        self.Schurn = d.Schurn
        self.Stmin = d.Stmin
        self.Stmax = d.Stmax
        self.Snoise = d.Snoise
        self.P_mean = getattr(d, "P_mean", -1.2)
        self.P_std = getattr(d, "P_std", 1.2)
        self.ro_train = getattr(d, "ro_train", d.ro)
        aw = getattr(d, "aweighting", None)
        if aw is not None and getattr(aw, "use_aweighting", False):
            raise NotImplementedError("A-weighted training loss (utils.training_utils.FIRFilter) is outside this path")

    def get_gamma(self, t):
        """diff_params/edm.py:38-53."""
        N = t.shape[0]
        gamma = torch.zeros(t.shape).to(t.device)
        sel = torch.logical_and(t > self.Stmin, t < self.Stmax)
        gamma[sel] = gamma[sel] + torch.min(torch.Tensor([self.Schurn / N, 2 ** (1 / 2) - 1]))
        return gamma

    def create_schedule(self, nb_steps):
        """diff_params/edm.py:55-64."""
        return self.create_schedule_from_initial_t(self.sigma_max, nb_steps)

    def create_schedule_from_initial_t(self, initial_t, nb_steps):
        """diff_params/edm.py:66-75: nb_steps+1 points, abscissa i/(nb_steps-1),
        the extrapolated last point is overwritten with 0."""
        i = torch.arange(0, nb_steps + 1)
        t = (initial_t ** (1 / self.ro) + i / (nb_steps - 1)
             * (self.sigma_min ** (1 / self.ro) - initial_t ** (1 / self.ro))) ** self.ro
        t[-1] = 0
        return t

    def sample_prior(self, shape, sigma):
        """diff_params/edm.py:98-106 (host generator, like the reference)."""
        return torch.randn(shape).to(sigma.device) * sigma

    def cskip(self, sigma):
        return self.sigma_data ** 2 * (sigma ** 2 + self.sigma_data ** 2) ** -1

    def cout(self, sigma):
        return sigma * self.sigma_data * (self.sigma_data ** 2 + sigma ** 2) ** (-0.5)

    def cin(self, sigma):
        return (self.sigma_data ** 2 + sigma ** 2) ** (-0.5)

    def cnoise(self, sigma):
        return (1 / 4) * torch.log(sigma)

    def denoiser(self, xn, net, sigma):
        """diff_params/edm.py:144-159."""
        if len(sigma.shape) == 1:
            sigma = sigma.unsqueeze(-1)
        return self.cskip(sigma) * xn + self.cout(sigma) * net(self.cin(sigma) * xn, self.cnoise(sigma))

    # -- training side (diff_params/edm.py:78-96, 141-142, 161-211) ----------------------------------
    def sample_ptrain(self, N):
        """Log-normal noise levels of Karras et al., clipped to [sigma_min, sigma_max] (:78-86)."""
        lnsigma = np.random.randn(N) * self.P_std + self.P_mean
        return np.clip(np.exp(lnsigma), self.sigma_min, self.sigma_max)

    def sample_ptrain_safe(self, N):
        """Noise levels drawn along the sampling schedule with exponent ``ro_train`` (:88-96)."""
        a = torch.rand(N)
        lo, hi = self.sigma_min ** (1 / self.ro_train), self.sigma_max ** (1 / self.ro_train)
        return (hi + a * (lo - hi)) ** self.ro_train

    def lambda_w(self, sigma):
        """:141-142."""
        return (sigma * self.sigma_data) ** (-2) * (self.sigma_data ** 2 + sigma ** 2)

    def prepare_train_preconditioning(self, x, sigma):
        """:161-174 -> (network input, regression target, noise conditioning)."""
        noise = self.sample_prior(x.shape, sigma)
        cskip, cout, cin, cnoise = self.cskip(sigma), self.cout(sigma), self.cin(sigma), self.cnoise(sigma)
        target = (1 / cout) * (x - cskip * (x + noise))
        return cin * (x + noise), target, cnoise

    def loss_fn(self, net, x):
        """:177-211 -> (squared error (B,T), sigma (B,1)); the DC/Nyquist bands the CQT network cannot
        represent are removed from the error when ``args.net.use_cqt_DC_correction`` is set (:194-199)."""
        sigma = self.sample_ptrain_safe(x.shape[0]).unsqueeze(-1).to(x.device)
        inp, target, cnoise = self.prepare_train_preconditioning(x, sigma)
        error = net(inp, cnoise) - target
        net_cfg = getattr(self.args, "net", None)
        if net_cfg is not None and getattr(net_cfg, "use_cqt_DC_correction", False) and hasattr(net, "CQTransform"):
            error = net.CQTransform.apply_hpf_DC(error)
        return error ** 2, sigma
