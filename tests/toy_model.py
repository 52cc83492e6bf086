"""A tiny deterministic stand-in denoiser shared by the golden generator and the
tests.  It has the interface the reference sampler needs from its model
(testing/blind_bwe_sampler.py:153-157): ``forward(x[B,T], cnoise[B,1]) -> [B,T]``
and an attribute ``CQTransform`` exposing ``apply_hpf_DC``."""
import torch


class _MeanRemoval:
    def apply_hpf_DC(self, x):
        return x - x.mean(dim=-1, keepdim=True)


class ToyDenoiser(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(1)
        self.taps = torch.nn.Parameter(0.2 * torch.randn(1, 1, 9, generator=g), requires_grad=False)
        self.CQTransform = _MeanRemoval()

    def forward(self, x, cnoise):
        z = torch.nn.functional.conv1d(x.unsqueeze(1), self.taps.to(x.dtype), padding=4).squeeze(1)
        return 0.1 * torch.tanh(4.0 * z) * (1.0 + 0.05 * cnoise) + 0.3 * z
