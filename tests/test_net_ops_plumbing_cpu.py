"""CPU: the autograd plumbing of babe_b200.net_ops (_ResLayer / _AddScale / _ConvSame: argument
handling, saved tensors, the flipped-weight cache, both convolution formulations, gradient arity)
with the four C-ABI wrappers replaced by plain torch restatements of what the kernels compute
(csrc/net_ops.cu header).  The kernels themselves are checked on the GPU (tests/test_net_ops_gpu.py)."""
import pytest
import torch

from conftest import rel_l2
from oracle import net_glue as og


@pytest.fixture
def torch_kernels(monkeypatch):
    from babe_b200 import net_ops

    def gn_stats(x, groups):
        n, c = x.shape[:2]
        xg = x.reshape(n, groups, -1).double()
        return torch.stack((xg.mean(-1), xg.std(-1)), -1), 1          # "partials": (mean, std) per group

    def _rs(x, part, groups, eps):
        n, c = x.shape[:2]
        std = part[..., 1].float()
        r = 1.0 / (std + eps)
        return (r.repeat_interleave(c // groups, 1), part[..., 0].float().repeat_interleave(c // groups, 1),
                std.repeat_interleave(c // groups, 1))

    def gn_film_gelu(x, part, S, gamma, aff, groups, eps):
        r, _, _ = _rs(x, part, groups, eps)
        s = r * gamma[None] * (aff + 1)
        return torch.nn.functional.gelu(x * s[:, :, None, None])

    def gate_residual(x0, v, gate, scale=net_ops.RSQRT2):
        g = 1.0 if gate is None else gate[:, :, None, None]
        return (v * g if x0 is None else x0 + v * g) * scale

    def gn_film_gelu_bwd(gh, x, gy, part, S, gamma, aff, groups, eps, res_scale=net_ops.RSQRT2):
        n, c = x.shape[:2]
        gc = c // groups
        r, mean, std = _rs(x, part, groups, eps)
        q = gamma[None] * (aff + 1)
        s = (r * q)[:, :, None, None]
        u = x * s
        cdf = 0.5 * (1 + torch.erf(u * 0.7071067811865476))
        gu = gh * (cdf + u * torch.exp(-0.5 * u * u) * 0.3989422804014327)
        cnt = gc * x.shape[2] * x.shape[3]
        Gr = (gu * x * q[:, :, None, None]).reshape(n, groups, -1).sum(-1).repeat_interleave(gc, 1)
        coef = (Gr * r * r / ((cnt - 1) * std))[:, :, None, None]
        return gy * res_scale + gu * s - coef * (x - mean[:, :, None, None])

    for name, fn in (("gn_stats", gn_stats), ("gn_film_gelu", gn_film_gelu), ("gate_residual", gate_residual),
                     ("gn_film_gelu_bwd", gn_film_gelu_bwd)):
        monkeypatch.setattr(net_ops, name, fn)
    yield net_ops
    net_ops._CONV_CHOICE = {}


@pytest.mark.parametrize("choice", [0, 1])
def test_res_layer_plumbing(torch_kernels, choice):
    net_ops = torch_kernels

    class Always(dict):
        def get(self, key, default=None):
            return choice
    net_ops._CONV_CHOICE = Always()
    g = torch.Generator().manual_seed(0)
    n, c = 2, 16
    x = (torch.randn(n, c, 7, 9, generator=g) * 1.3 + 0.1).requires_grad_(True)
    gamma = torch.nn.Parameter(1 + 0.2 * torch.randn(1, c, 1, 1, generator=g), requires_grad=False)
    w = torch.nn.Parameter(torch.randn(c, c, 5, 3, generator=g) * 0.1, requires_grad=False)
    aff, gate = 0.3 * torch.randn(1, c, generator=g), torch.randn(n, c, generator=g)     # aff broadcast over n
    gy = torch.randn(n, c, 7, 9, generator=g)
    y = net_ops.add_scale(net_ops.res_layer(x, gamma, aff, gate, w, (2, 1), 8, 1e-7), x)
    gx, = torch.autograd.grad(y, x, gy)
    yr = og.add_scale(og.res_layer(x, gamma.reshape(-1), aff.expand(n, c), gate, w, (2, 1)), x)
    gr, = torch.autograd.grad(yr, x, gy)
    assert rel_l2(y.detach(), yr.detach()) < 1e-5 and rel_l2(gx, gr) < 1e-5
    assert (id(w) in net_ops._FLIPPED) == (choice == 1)            # flipped copy cached per weight object


def test_usable_gates():
    from babe_b200 import net_ops
    x = torch.zeros(1, 8, 2, 2)
    assert not net_ops.usable(x)                                   # CPU tensors never take the fused path
    p = torch.nn.Parameter(torch.zeros(3))
    assert not net_ops._frozen(p) and net_ops._frozen(p.detach())
    with torch.no_grad():
        assert net_ops._frozen(p)
