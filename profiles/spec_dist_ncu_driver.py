"""ncu driver: one launch of each spectrogram-loss kernel at B = 64 x T = 184184 (F = 2049, 88 frames)."""
import sys
sys.path.insert(0, ".")
import torch
from babe_b200 import ops
dev = torch.device("cuda")
B, F, M = 64, 2049, 88
X = torch.randn(B, F, M, 2, device=dev)
R = torch.randn(B, F, M, 2, device=dev)
w = torch.linspace(0, 1, F, device=dev)
coef = torch.ones(1, device=dev)
ops.spec_dist_stats(X, R, w, 0); ops.spec_dist_stats(X, R, w, 2); ops.spec_mag_stats(X, R, None, w)
ops.spec_dist_grad(X, R, w, coef, 0); ops.spec_dist_grad(X, R, w, coef, 2); ops.spec_mag_grad(X, R, None, w, coef)
torch.cuda.synchronize()
print("ok")
