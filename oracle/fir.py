"""Oracle: classical FIR observation model (CPU, torch).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.  Restates
``apply_low_pass_firwin`` (utils/bandwidth_extension.py:76-95 of eloimoliner/BABE) and
``BlindSampler.apply_FIR_filter`` (testing/blind_bwe_sampler.py:211-218) as an explicit
correlation so that the 'same' padding convention is visible; pinned by
tests/golden/fir.npz, which the reference's own function produced.
"""
import torch


def apply_fir_same(y, taps):
    """y (B,T), taps (L,) -> (B,T):  out[n] = sum_k taps[k] y[n + k - (L-1)//2], zero outside."""
    taps = taps.reshape(-1).to(y.dtype)
    L = taps.numel()
    pl = (L - 1) // 2
    yp = torch.nn.functional.pad(y, (pl, L - 1 - pl))
    return (yp.unfold(1, L, 1) * taps).sum(-1)


def apply_fir_same_adjoint(g, taps):
    """Transpose of ``apply_fir_same`` wrt y."""
    taps = taps.reshape(-1).to(g.dtype)
    L = taps.numel()
    pl = (L - 1) // 2
    gp = torch.nn.functional.pad(g, (L - 1 - pl, pl))
    return (gp.unfold(1, L, 1) * taps.flip(0)).sum(-1)
