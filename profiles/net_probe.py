"""Tiny driver for ncu: one fused residual layer (forward + backward) and the x2 resamplers at a
shape of the sampler benchmark (second U-Net level at 8 chains: 8 x 96 x 128 x 1024)."""
import sys
sys.path.insert(0, ".")
import torch
from babe_b200 import denoiser, net_ops
torch.backends.cudnn.allow_tf32 = True
N, C, Fd, T = 8, 96, 128, 1024
x = torch.randn(N, C, Fd, T, device="cuda", requires_grad=True)
gamma = torch.ones(1, C, 1, 1, device="cuda")
aff = torch.randn(N, C, device="cuda") * 0.1
gate = torch.randn(N, C, device="cuda") * 0.1
w = torch.randn(C, C, 5, 3, device="cuda") * 0.02
for _ in range(3):
    y = net_ops.res_layer(x, gamma, aff, gate, w, (2, 1), 8, 1e-7)
    gx, = torch.autograd.grad(y, x, y)
    d = net_ops.resample2(x, denoiser._CUBIC, False)
    u = net_ops.resample2(d, denoiser._CUBIC, True)
    gd, = torch.autograd.grad(u, x, u)
torch.cuda.synchronize()
print("ok", float(gx.abs().mean()), float(gd.abs().mean()))
