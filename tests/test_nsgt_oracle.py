"""CPU: invariants of the NSGT oracle (the CQT has no upstream oracle -- parity
unpinned -- so the restatement is validated through the transform's own
properties)."""
import pytest
import torch

from conftest import rel_l2
from oracle.nsgt import NSGT


@pytest.mark.parametrize("numocts,binsoct,fs,Ls", [(7, 64, 22050, 184184), (4, 12, 22050, 8192),
                                                   (5, 24, 44100, 30030)])
def test_structure_and_invariants(numocts, binsoct, fs, Ls):
    t = NSGT(numocts, binsoct, fs, Ls, ("kaiser", 1))
    # contract: T doubling, painless condition
    for o in range(numocts - 1):
        assert t.M[o + 1] == 2 * t.M[o]
    for j, L in enumerate(t.Lg):
        assert L <= t.M[j // binsoct]
    if Ls == 184184:
        assert t.M == [32, 64, 128, 256, 512, 1024, 2048]
    # frame operator diagonal is strictly positive, Hhpf in [0,1], kills DC / Nyquist
    assert float(t.D.min()) > 0
    assert float(t.Hhpf.min()) >= 0 and float(t.Hhpf.max()) <= 1 + 1e-12
    assert float(t.Hhpf[0]) < 1e-12 and float(t.Hhpf[-1]) < 1 - 1e-3   # DC removed, Nyquist shared with the Nyquist band
    mid = t.Hhpf[t.p[0] + t.Lg[0]: t.p[-1] - t.Lg[-1]]
    assert float((mid - 1).abs().max()) < 1e-12          # partition of unity away from the edges
    B = 2
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, Ls, generator=g, dtype=torch.float64)
    c = t.fwd(x)
    assert len(c) == numocts and all(c[o].shape == (B, binsoct, t.M[o]) for o in range(numocts))
    # the vectorised statement equals the band-by-band definition
    if Ls <= 30030:
        cl = t.fwd_loop(x)
        assert all(rel_l2(torch.view_as_real(a), torch.view_as_real(b)) < 1e-13 for a, b in zip(c, cl))
        assert rel_l2(t.bwd(c), t.bwd_loop(c)) < 1e-13
    # perfect reconstruction up to the DC/Nyquist high-pass
    assert rel_l2(t.bwd(c), t.apply_hpf_DC(x)) < 1e-12
    # linearity
    x2 = torch.randn(B, Ls, generator=g, dtype=torch.float64)
    c2 = t.fwd(0.3 * x - 1.7 * x2)
    cc = t.fwd(x2)
    assert all(rel_l2(torch.view_as_real(c2[o]), torch.view_as_real(0.3 * c[o] - 1.7 * cc[o])) < 1e-12
               for o in range(numocts))


def test_adjoints_and_autograd():
    t = NSGT(4, 12, 22050, 8192, ("kaiser", 1))
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 8192, generator=g, dtype=torch.float64, requires_grad=True)
    c = t.fwd(x)
    cot = [torch.randn(ci.shape, generator=g, dtype=torch.float64) + 1j * torch.randn(ci.shape, generator=g, dtype=torch.float64)
           for ci in c]
    # <fwd x, cot>_R differentiated = Re(fwd^H cot); check with a directional derivative
    loss = sum((torch.view_as_real(ci) * torch.view_as_real(co)).sum() for ci, co in zip(c, cot))
    (gx,) = torch.autograd.grad(loss, x)
    d = torch.randn(1, 8192, generator=g, dtype=torch.float64)
    cd = t.fwd(d)
    lhs = sum((torch.view_as_real(ci) * torch.view_as_real(co)).sum() for ci, co in zip(cd, cot))
    assert abs(float(lhs - (gx * d).sum())) < 1e-9 * abs(float(lhs))


def test_tone_localisation():
    t = NSGT(7, 64, 22050, 184184, ("kaiser", 1))
    j = 300
    f = t.p[j] * 22050 / 184184
    n = torch.arange(184184, dtype=torch.float64)
    x = torch.sin(2 * torch.pi * f * n / 22050)[None]
    c = t.fwd(x)
    e = torch.cat([ci.abs().pow(2).sum(-1)[0] / t.M[o] for o, ci in enumerate(c)])
    assert int(e.argmax()) == j
