"""Karras EDM schedule and preconditioning used by the blind sampler.

Host-side scalar mathematics mirroring ``diff_params/edm.py`` (EDM class) of
eloimoliner/BABE so that ``args.diff_params.callable`` can point here when the
reference tree is not importable (e.g. on the GPU box).  Same attribute and
method names; the training-only members are omitted.
"""
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
