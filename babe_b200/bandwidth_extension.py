"""Drop-in for the FIR part of ``utils.bandwidth_extension`` of eloimoliner/BABE
(SURVEY 8f-4): the classical low-pass observation model of the non-blind baseline.
Filter design stays scipy (host, once per filter); applying it runs on the CUDA
overlap-save kernel (``babe_fir_filter``) and is differentiable wrt the signal."""
import torch

from . import ops


def get_FIR_lowpass(order, fc, beta, sr):
    """utils/bandwidth_extension.py:59-74 -> (1,1,order) taps."""
    import scipy.signal
    B = scipy.signal.firwin(numtaps=order, cutoff=fc, width=beta, window="kaiser", fs=sr)
    return torch.FloatTensor(B).unsqueeze(0).unsqueeze(0)


def get_FIR_high_pass(order, fc, beta, sr):
    """utils/bandwidth_extension.py:43-58 -> (1,1,order-1) taps."""
    import scipy.signal
    B = scipy.signal.firwin(numtaps=order - 1, cutoff=fc, width=beta, window="kaiser", fs=sr, pass_zero="highpass")
    return torch.FloatTensor(B).unsqueeze(0).unsqueeze(0)


class _FirOp(torch.autograd.Function):
    """conv1d(padding='same') with fixed taps; linear in y, so the backward is the same
    Function with the adjoint flag flipped."""

    @staticmethod
    def forward(ctx, y, taps, adjoint):
        ctx.taps, ctx.adjoint = taps, adjoint
        return ops.fir_filter(y, taps, adjoint=adjoint)

    @staticmethod
    def backward(ctx, g):
        return _FirOp.apply(g.contiguous(), ctx.taps, not ctx.adjoint), None, None


def apply_low_pass_firwin(y, filter):
    """utils/bandwidth_extension.py:76-95: y (B,T), filter (1,1,L) -> (B,T)."""
    return _FirOp.apply(y, filter.detach(), False)
