"""Tiny driver for ncu: the CQT operators alone at the sampler's batch."""
import sys
sys.path.insert(0, ".")
import torch
from cqt_nsgt_pytorch import CQT_nsgt
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cq = CQT_nsgt(7, 64, mode="oct", window=("kaiser", 1), fs=22050, audio_len=184184, device="cuda")
x = torch.randn(B, 1, 184184, device="cuda")
for _ in range(3):
    c = cq.fwd(x)
    y = cq.bwd(c)
    z = cq.apply_hpf_DC(x.squeeze(1))
torch.cuda.synchronize()
print("ok")
