"""Tiny driver for ncu: apply_filter (forward, adjoint) and the fit statistics at the operator-sweep
shape B = 512, T = 2^17, NFFT = 4096."""
import sys
sys.path.insert(0, ".")
import torch
from babe_b200 import ops
B, T, NFFT = 512, 1 << 17, 4096
dev = torch.device("cuda")
x = torch.randn(B, T, device=dev) * 0.063
y = torch.randn(B, T, device=dev) * 0.063
out = torch.empty_like(x)
f = torch.fft.rfftfreq(NFFT, d=1 / 22050).to(dev)
fc = torch.tensor([300.0, 600.0, 1000.0, 3000.0, 6000.0], device=dev)
A = torch.tensor([-10.0, -15.0, -20.0, -30.0, -40.0], device=dev)
for _ in range(3):
    ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, out=out)
    ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, adjoint=True, out=out)
    ops.stft_stats(x, y, NFFT)
torch.cuda.synchronize()
print("ok")
