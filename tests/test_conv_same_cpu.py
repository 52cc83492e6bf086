"""CPU: the two formulations of the stride-1 "same" convolution used by babe_b200.net_ops
(cuDNN fprop vs dgrad of the flipped-transposed weight) are the same linear map, forward and
backward, and the flipped-weight cache never serves a stale copy."""
import gc

import torch
import torch.nn.functional as F

from conftest import rel_l2


def _force(choice):
    """Make conv_same use formulation `choice` for every key without timing anything."""
    from babe_b200 import net_ops

    class Always(dict):
        def get(self, key, default=None):
            return choice
    net_ops._CONV_CHOICE = Always()
    net_ops.AUTOTUNE_CONV = True
    return net_ops


def test_both_formulations_agree():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 6, 9, 11, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.randn(4, 6, 5, 3, generator=g, dtype=torch.float64)
    gy = torch.randn(2, 4, 9, 11, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, padding="same", dilation=(2, 1))
    gref, = torch.autograd.grad(ref, x, gy)
    try:
        for choice in (0, 1):
            net_ops = _force(choice)
            y = net_ops.conv_frozen(x, w, (2, 1))
            gx, = torch.autograd.grad(y, x, gy)
            assert rel_l2(y, ref) < 1e-12 and rel_l2(gx, gref) < 1e-12
    finally:
        net_ops._CONV_CHOICE = {}


def test_flipped_cache_is_per_object_and_versioned():
    from babe_b200 import net_ops
    w = torch.randn(3, 2, 5, 3)
    f1 = net_ops._flipped(w)
    assert net_ops._flipped(w) is f1                           # cached
    assert torch.equal(f1, w.transpose(0, 1).flip(2, 3))
    w.mul_(2)                                                  # in-place update bumps the version
    f2 = net_ops._flipped(w)
    assert f2 is not f1 and torch.equal(f2, w.transpose(0, 1).flip(2, 3))
    k = id(w)
    del w, f1, f2
    gc.collect()
    assert k not in net_ops._FLIPPED                           # entry dropped with the object
    # a new tensor (possibly at the recycled address / id) gets its own copy
    w2 = torch.randn(3, 2, 5, 3)
    assert torch.equal(net_ops._flipped(w2), w2.transpose(0, 1).flip(2, 3))
