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


# ---------------------------------------------------------------------------
# The rest of the module's surface (design helpers, dispatcher).  The IIR / biquad / polyphase
# observation models are sequential recursions used once to synthesise y; they are NOT on the
# hand-written path and are delegated to torchaudio exactly like the reference
# does, so that a tester importing this module in place of ``utils.bandwidth_extension`` finds
# every name it uses.
# ---------------------------------------------------------------------------
def get_cheby1_ba(order, ripple, hi):
    """utils/bandwidth_extension.py:169-178: Chebyshev-I lowpass (b, a) from scipy."""
    import scipy.signal
    return scipy.signal.cheby1(order, ripple, hi, btype="lowpass", output="ba")


def design_biquad_lpf(fc, fs, Q):
    """utils/bandwidth_extension.py:180-199: RBJ cookbook lowpass, (b0, b1, b2, a0, a1, a2) as 0-d tensors."""
    import math
    w0 = torch.as_tensor(2 * math.pi * fc / fs, dtype=torch.float32)
    c, alpha = torch.cos(w0), torch.sin(w0) / 2 / Q
    half = (1 - c) / 2
    return half, 1 - c, half, 1 + alpha, -2 * c, 1 - alpha


def prepare_filter(args, sample_rate):
    """utils/bandwidth_extension.py:7-40: the filter object for ``tester.bandwidth_extension.filter.type``."""
    cfg = args.tester.bandwidth_extension.filter
    kind = cfg.type
    if kind == "firwin":
        return get_FIR_lowpass(cfg.order, cfg.fc, cfg.beta, sample_rate)
    if kind == "firwin_hpf":
        return get_FIR_high_pass(cfg.order, cfg.fc, cfg.beta, sample_rate)
    if kind == "cheby1":
        return get_cheby1_ba(cfg.order, cfg.ripple, 2 * cfg.fc / sample_rate)
    if kind == "biquad":
        return design_biquad_lpf(cfg.fc, sample_rate, cfg.biquad.Q)
    if kind == "resample":
        return sample_rate / cfg.resample.fs
    if kind == "decimate":
        factor = int(args.tester.bandwidth_extension.decimate.factor)
        cfg.resample.fs = int(sample_rate / factor)          # side effect of the reference (:31)
        return factor
    raise NotImplementedError(kind)


def apply_decimate(y, factor):
    """utils/bandwidth_extension.py:97-108: naive decimation y[..., 0:-1:factor]."""
    return y[..., 0:-1:factor]


def apply_resample(y, factor):
    """utils/bandwidth_extension.py:110-118 (torchaudio polyphase resampler; library call)."""
    import torchaudio
    N = 100
    return torchaudio.functional.resample(y, orig_freq=int(factor * N), new_freq=N)


def apply_low_pass_biquad(y, filter):
    """utils/bandwidth_extension.py:120-137 (torchaudio biquad; library call)."""
    import torchaudio
    b0, b1, b2, a0, a1, a2 = (torch.as_tensor(v, dtype=torch.float32).to(y.device) for v in filter)
    return torchaudio.functional.biquad(y, b0, b1, b2, a0, a1, a2)


def apply_low_pass_IIR(y, filter):
    """utils/bandwidth_extension.py:138-143 (torchaudio lfilter; library call).  The reference passes
    (a, b) swapped into lfilter's (a_coeffs, b_coeffs) slots as written there: kept."""
    import torchaudio
    b, a = filter
    b = torch.as_tensor(b, dtype=torch.float32).to(y.device)
    a = torch.as_tensor(a, dtype=torch.float32).to(y.device)
    return torchaudio.functional.lfilter(y, a, b, clamp=False)


def apply_low_pass(y, filter, type):
    """utils/bandwidth_extension.py:145-167: dispatch on the filter type (None for unknown types, like
    the reference)."""
    if type in ("firwin", "firwin_hpf"):
        return apply_low_pass_firwin(y, filter)
    if type == "cheby1":
        return apply_low_pass_IIR(y, filter)
    if type == "biquad":
        return apply_low_pass_biquad(y, filter)
    if type == "resample":
        return apply_resample(y, filter)
    if type == "decimate":
        return apply_decimate(y, filter)
    return None
