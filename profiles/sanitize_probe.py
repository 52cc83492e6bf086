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
torch.cuda.synchronize()
print("ok", float(y.abs().sum()), float(yp.abs().sum()), float(xr.abs().sum()), q.tolist()[0][:2])
