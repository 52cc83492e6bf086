"""Tiny driver for ncu: one statistics pass + filter fits (100 iterations, K=5)."""
import sys
sys.path.insert(0, ".")
import torch
from babe_b200 import ops, sampler
dev = torch.device("cuda")
x = torch.randn(8, 184184, device=dev) * 0.063
y = torch.randn(8, 184184, device=dev) * 0.063
fit = sampler.FilterFit(nfft=4096, sample_rate=22050, device=dev)
abc = ops.stft_stats(x, y, 4096)
p = torch.tensor([[280.0, 285, 290, 295, 300], [-15.0, -17, -20, -25, -30]], device=dev)
for _ in range(3):
    q = fit(x, y, p.clone(), abc=abc)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10):
    q = fit(x, y, p.clone(), abc=abc)
e.record()
torch.cuda.synchronize()
print("fit ms", s.elapsed_time(e) / 10, q.tolist())
