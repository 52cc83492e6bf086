"""Small driver for compute-sanitizer (memcheck / racecheck): every CQT kernel once at B = 2, both layouts, plus the fit."""
import sys
sys.path.insert(0, ".")
import torch
from cqt_nsgt_pytorch import CQT_nsgt
from babe_b200 import ops, sampler
dev = torch.device("cuda")
cq = CQT_nsgt(7, 64, mode="oct", window=("kaiser", 1), fs=22050, audio_len=184184, device=dev)
x = torch.randn(2, 184184, device=dev) * 0.063
c = cq.fwd(x.unsqueeze(1)); y = cq.bwd(c); z = cq.apply_hpf_DC(x)
cp = cq.fwd_planar(x); yp = cq.bwd_planar(cp)
X = cq.rfft(x); xr = cq.irfft(X)
fit = sampler.FilterFit(nfft=4096, sample_rate=22050, device=dev, max_iter=3)
p = torch.tensor([[280.0, 285, 290, 295, 300], [-15.0, -17, -20, -25, -30]], device=dev)
q = fit(x, z, p.clone())
# the fused NFFT-4096 operators (TMA-staged tiles), every epilogue variant, and the other lengths' CQT plans
f = torch.fft.rfftfreq(4096, d=1 / 22050).to(dev)
fc, A = torch.tensor([300.0, 1000.0, 4000.0], device=dev), torch.tensor([-10.0, -20.0, -40.0], device=dev)
xs = torch.randn(3, 60000, device=dev) * 0.063
o1 = ops.apply_filter(xs, 4096, freqs=f, fc=fc, A=A)
o2 = ops.apply_filter(xs, 4096, freqs=f, fc=fc, A=A, adjoint=True)
st = ops.stft_stats(xs, o1, 4096)
from babe_b200 import blind_bwe_utils as bu
xg = xs.clone().requires_grad_(True)
H = bu.design_filter(fc, A, f)
(bu.apply_filter(xg, H, 4096) - o1).pow(2).sum().backward()
for (no, bo, fs, L) in ((7, 64, 22050, 132300), (8, 96, 44100, 485100), (8, 96, 44100, 368368), (4, 12, 22050, 8192)):
    cq2 = CQT_nsgt(no, bo, mode="oct", window=("kaiser", 1), fs=fs, audio_len=L, device=dev)
    x2 = torch.randn(1, L, device=dev) * 0.063
    r2 = cq2.bwd(cq2.fwd(x2.unsqueeze(1))); h2 = cq2.apply_hpf_DC(x2)
torch.cuda.synchronize()
print("ok", float(y.abs().sum()), float(yp.abs().sum()), float(xr.abs().sum()), q.tolist()[0][:2])
