"""CPU: the oracle restatement of the denoiser layer glue (oracle/net_glue.py) against golden
vectors produced by the UNMODIFIED reference modules (tests/golden/make_golden_net.py)."""
import math

import torch

from conftest import rel_l2
from oracle import net_glue as og


def _t(g, k, dtype=torch.float32):
    return torch.from_numpy(g[k]).to(dtype)


def _block(g, x, dtype):
    y = x
    for i in range(2):
        y = og.res_layer(y, _t(g, f"blk_gamma{i}", dtype), _t(g, f"blk_aff{i}", dtype), _t(g, f"blk_gate{i}", dtype),
                         _t(g, f"blk_w{i}", dtype), (2 ** i, 1))
    return og.add_scale(y, x)                                   # networks/cqtdiff+.py:487 (identity res_conv)


def test_group_norm_matches_reference(golden):
    g = golden("net_glue.npz")
    y = og.bias_free_group_norm(_t(g, "blk_x"), _t(g, "blk_gamma0"), 8)
    assert rel_l2(y, g["norm_y"]) < 1e-6


def test_resnet_block_matches_reference(golden):
    g = golden("net_glue.npz")
    for dtype, tol in ((torch.float32, 1e-6), (torch.float64, 1e-6)):
        x = _t(g, "blk_x", dtype).requires_grad_(True)
        y = _block(g, x, dtype)
        gx, = torch.autograd.grad(y, x, _t(g, "blk_gy", dtype))
        assert rel_l2(y.detach(), g["blk_y"]) < tol
        assert rel_l2(gx, g["blk_gx"]) < 10 * tol


def test_resamplers_match_reference(golden):
    g = golden("net_glue.npz")
    for name, up in (("down", False), ("up", True)):
        x = _t(g, f"rs_{name}_x").requires_grad_(True)
        y = og.resample2(x, up)
        gx, = torch.autograd.grad(y, x, _t(g, f"rs_{name}_gy"))
        assert y.shape == g[f"rs_{name}_y"].shape
        assert rel_l2(y.detach(), g[f"rs_{name}_y"]) < 1e-6
        assert rel_l2(gx, g[f"rs_{name}_gx"]) < 1e-6


def test_resampler_is_half_band():
    """Size-independent property: the cubic taps sum to 1 per phase pair (DC gain 1 down, 2 x 0.5 up)."""
    assert math.isclose(sum(og.CUBIC), 1.0, abs_tol=1e-12)
    x = torch.ones(1, 1, 1, 64)
    assert torch.allclose(og.resample2(x, False), torch.ones(1, 1, 1, 32), atol=1e-6)
    assert torch.allclose(og.resample2(x, True), torch.full((1, 1, 1, 128), 0.5), atol=1e-6)
