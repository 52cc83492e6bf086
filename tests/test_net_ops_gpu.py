"""GPU: the fused residual-layer glue (csrc/net_ops.cu) against the composite PyTorch
expression of networks/cqtdiff+.py:470-482 (fp32 reference of the same op), forward and
gradient wrt the activations."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu


def _composite(x, gamma, aff, gate, weight, dilation, groups, eps):
    n, c, f, t = x.shape
    xg = x.reshape(n, groups, -1)
    xg = xg / (xg.std(-1, keepdim=True) + eps)
    h = xg.reshape(n, c, f, t) * gamma
    h = h * (aff[:, :, None, None] + 1)
    v = F.conv2d(F.gelu(h), weight, padding="same", dilation=dilation)
    return (x + v * gate[:, :, None, None]) / (2 ** 0.5)


@pytest.mark.parametrize("shape,dil", [((2, 64, 64, 16), (1, 1)), ((3, 96, 20, 37), (4, 1)),
                                       ((1, 256, 8, 130), (2, 1)), ((2, 16, 33, 1025), (1, 1))])
def test_res_layer_matches_composite(shape, dil):
    from babe_b200 import build, net_ops
    build.build()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    n, c = shape[:2]
    x = (torch.randn(shape, device="cuda", generator=g) * 1.7 + 0.3).requires_grad_(True)
    gamma = 1 + 0.2 * torch.randn(1, c, 1, 1, device="cuda", generator=g)
    aff = 0.5 * torch.randn(n, c, device="cuda", generator=g)
    gate = torch.randn(n, c, device="cuda", generator=g)
    w = torch.randn(c, c, 5, 3, device="cuda", generator=g) / (c * 15) ** 0.5
    gy = torch.randn(shape, device="cuda", generator=g)
    y_ref = _composite(x, gamma, aff, gate, w, dil, 8, 1e-7)
    gx_ref, = torch.autograd.grad(y_ref, x, gy)
    y = net_ops.res_layer(x, gamma, aff, gate, w, dil, 8, 1e-7)
    gx, = torch.autograd.grad(y, x, gy)
    assert rel_l2(y, y_ref) < 1e-5
    assert rel_l2(gx, gx_ref) < 1e-5
    # bitwise reproducible (the glue kernels always; cuDNN's dgrad kernels, which the convolution
    # autotuner may pick for the forward pass, only in cuDNN's deterministic mode)
    prev = torch.backends.cudnn.deterministic
    torch.backends.cudnn.deterministic = True
    try:
        y1 = net_ops.res_layer(x, gamma, aff, gate, w, dil, 8, 1e-7)
        y2 = net_ops.res_layer(x, gamma, aff, gate, w, dil, 8, 1e-7)
    finally:
        torch.backends.cudnn.deterministic = prev
    assert torch.equal(y1, y2)


def test_add_scale():
    from babe_b200 import build, net_ops
    build.build()
    a = torch.randn(2, 5, 7, 13, device="cuda", requires_grad=True)
    b = torch.randn(2, 5, 7, 13, device="cuda", requires_grad=True)
    y = net_ops.add_scale(a, b)
    ga, gb = torch.autograd.grad(y, (a, b), torch.ones_like(y))
    assert rel_l2(y, (a + b) / 2 ** 0.5) < 1e-6
    assert rel_l2(ga, torch.full_like(a, 2 ** -0.5)) < 1e-6 and torch.equal(ga, gb)


@pytest.mark.parametrize("T", [8, 16, 50, 1024])
@pytest.mark.parametrize("up", [False, True])
def test_resample_matches_reference_formulation(T, up):
    """networks/cqtdiff+.py:522-580 restated literally (dense diagonal weight, conv1d / conv_transpose1d)."""
    from babe_b200 import build, denoiser, net_ops
    build.build()
    torch.backends.cudnn.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(T)
    x = torch.randn(2, 3, 5, T, device="cuda", generator=g, requires_grad=True)
    k = torch.tensor(denoiser._CUBIC, device="cuda")
    pad = len(denoiser._CUBIC) // 2 - 1
    xr = x.view(-1, x.shape[-2], x.shape[-1])
    w = xr.new_zeros(xr.shape[1], xr.shape[1], k.numel())
    idx = torch.arange(xr.shape[1], device="cuda")
    w[idx, idx] = k
    if up:
        ref = F.conv_transpose1d(F.pad(xr, ((pad + 1) // 2,) * 2, "reflect"), w, stride=2, padding=pad * 2 + 1)
    else:
        ref = F.conv1d(F.pad(xr, (pad,) * 2, "reflect"), w, stride=2)
    ref = ref.view(2, 3, 5, -1)
    y = net_ops.resample2(x, denoiser._CUBIC, up)
    assert y.shape == ref.shape
    assert rel_l2(y, ref) < 1e-6
    gy = torch.randn(ref.shape, device="cuda", generator=g)
    g_ref, = torch.autograd.grad(ref, x, gy)
    g_new, = torch.autograd.grad(y, x, gy)
    assert rel_l2(g_new, g_ref) < 1e-6


def test_denoiser_fused_matches_composite():
    """Whole CQTDiff+ body (small configuration): fused glue vs composite PyTorch, frozen parameters."""
    from babe_b200 import build, denoiser, sampler
    build.build()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    args = sampler.make_args(sample_rate=22050, audio_len=16384)
    args.network.Ns = [16, 16, 24, 24, 32, 32, 32]
    args.network.num_dils = [1, 2, 2, 2, 2, 2, 2]
    torch.manual_seed(0)
    net = denoiser.CQTDiffPlus(args, "cuda").cuda()
    # the gates are initialised ~0 (init_zero); give them weight so every branch matters
    with torch.no_grad():
        for name, p in net.named_parameters():
            if ".gate." in name:
                p.mul_(3e6)
    net.requires_grad_(False)
    x = (torch.randn(2, 16384, device="cuda") * 0.1).requires_grad_(True)
    sigma = torch.tensor([[0.5], [2.0]], device="cuda")
    outs = []
    for fused in (False, True):
        denoiser.FUSED = fused
        y = net(x, sigma)
        gx, = torch.autograd.grad(y.square().sum(), x)
        outs.append((y.detach(), gx))
    denoiser.FUSED = True
    assert rel_l2(outs[1][0], outs[0][0]) < 2e-5
    assert rel_l2(outs[1][1], outs[0][1]) < 1e-4


def test_glue_matches_reference_golden(golden):
    """The CUDA glue through the C ABI against vectors produced by the reference's own ResnetBlock /
    UpDownResample (tests/golden/make_golden_net.py)."""
    from babe_b200 import build, net_ops
    from oracle.net_glue import CUBIC
    build.build()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    g = golden("net_glue.npz")
    c = lambda k: torch.from_numpy(g[k]).cuda()
    x = c("blk_x").requires_grad_(True)
    y = x
    for i in range(2):
        y = net_ops.res_layer(y, c(f"blk_gamma{i}"), c(f"blk_aff{i}"), c(f"blk_gate{i}"), c(f"blk_w{i}"),
                              (2 ** i, 1), 8, 1e-7)
    y = net_ops.add_scale(y, x)
    gx, = torch.autograd.grad(y, x, c("blk_gy"))
    assert rel_l2(y.detach().cpu(), g["blk_y"]) < 1e-5
    assert rel_l2(gx.cpu(), g["blk_gx"]) < 1e-5
    for name, up in (("down", False), ("up", True)):
        xr = c(f"rs_{name}_x").requires_grad_(True)
        yr = net_ops.resample2(xr, CUBIC, up)
        gr, = torch.autograd.grad(yr, xr, c(f"rs_{name}_gy"))
        assert rel_l2(yr.detach().cpu(), g[f"rs_{name}_y"]) < 1e-6
        assert rel_l2(gr.cpu(), g[f"rs_{name}_gx"]) < 1e-6
