"""Tiny driver for ncu / quick timing: the CQT operators alone at batch B (default: the sampler's 8)."""
import sys
sys.path.insert(0, ".")
import torch
from cqt_nsgt_pytorch import CQT_nsgt
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cq = CQT_nsgt(7, 64, mode="oct", window=("kaiser", 1), fs=22050, audio_len=184184, device="cuda")
x = torch.randn(B, 1, 184184, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(3):
    c = cq.fwd(x)
    y = cq.bwd(c)
    z = cq.apply_hpf_DC(x.squeeze(1))
torch.cuda.synchronize()
C = sum(v.numel() for v in c) // B
for name, fn, nbytes in (("fwd", lambda: cq.fwd(x), 4 * B * 184184 + 8 * B * C),
                         ("bwd", lambda: cq.bwd(c), 4 * B * 184184 + 8 * B * C),
                         ("hpf", lambda: cq.apply_hpf_DC(x.squeeze(1)), 8 * B * 184184)):
    ts = []
    for _ in range(10):
        flush.fill_(1)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    print(f"B={B} {name}: {ms:.4f} ms  {nbytes / ms / 1e6:.1f} GB/s")
