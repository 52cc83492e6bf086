"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the CQTDiff+ layer glue that csrc/net_ops.cu
fuses.  Plain PyTorch on whatever device the inputs live on (tests use the CPU, fp32
or fp64); pinned by tests/golden/net_glue.npz, which tests/golden/make_golden_net.py produced
with the UNMODIFIED reference modules (ResnetBlock, BiasFreeGroupNorm, UpDownResample of
networks/cqtdiff+.py).  Never imported by the product path.
"""
import torch
import torch.nn.functional as F

CUBIC = [-0.01171875, -0.03515625, 0.11328125, 0.43359375,
         0.43359375, 0.11328125, -0.03515625, -0.01171875]      # networks/cqtdiff+.py:512-514


def bias_free_group_norm(x, gamma, groups, eps=1e-7):
    """networks/cqtdiff+.py:137-163: x / (unbiased std over (channels of the group, F, T) + eps) * gamma."""
    n, c, f, t = x.shape
    xg = x.reshape(n, groups, -1)
    xg = xg / (xg.std(-1, keepdim=True) + eps)
    return xg.reshape(n, c, f, t) * gamma.reshape(1, c, 1, 1)


def res_layer(x, gamma, aff, gate, weight, dilation, groups=8, eps=1e-7):
    """One iteration of the loop at networks/cqtdiff+.py:470-482 (aff = affine(sigma), gate = gate(sigma))."""
    h = bias_free_group_norm(x, gamma, groups, eps) * (aff[:, :, None, None] + 1)
    v = F.conv2d(F.gelu(h), weight, padding="same", dilation=dilation)
    return (x + v * gate[:, :, None, None]) / (2 ** 0.5)


def add_scale(a, b):
    """networks/cqtdiff+.py:487 and :792, :813: (a + b) / sqrt(2)."""
    return (a + b) / (2 ** 0.5)


def resample2(x, up, taps=CUBIC):
    """networks/cqtdiff+.py:522-580, mode "T": every (n, c, f) row filtered independently (the
    reference's dense C x C x L weight is non-zero on the diagonal only, :564-570)."""
    k = torch.tensor(taps, dtype=x.dtype, device=x.device)
    pad = len(taps) // 2 - 1
    rows = x.reshape(-1, 1, x.shape[-1])
    if up:
        rows = F.pad(rows, ((pad + 1) // 2,) * 2, "reflect")
        y = F.conv_transpose1d(rows, k.reshape(1, 1, -1), stride=2, padding=pad * 2 + 1)
    else:
        rows = F.pad(rows, (pad,) * 2, "reflect")
        y = F.conv1d(rows, k.reshape(1, 1, -1), stride=2)
    return y.reshape(*x.shape[:-1], y.shape[-1])
