"""Driver for the round-2 ncu captures (one launch of every hot kernel after warm-up):
    ncu --set full --clock-control none --import-source on -k regex:'k_filter_fused|k_stats_fused|k_fir_fused|k_fit_params2|k_cqt|k_fft' \
        -o gpurun_out/r02_kernels python profiles/r02_ncu_driver.py
"""
import sys
sys.path.insert(0, ".")
import torch
from babe_b200 import ops, sampler, bandwidth_extension as bwe
from cqt_nsgt_pytorch import CQT_nsgt

NFFT, SR = 4096, 22050
dev = torch.device("cuda")
B, T = 512, 1 << 17
x = torch.randn(B, T, device=dev) * 0.063
y = torch.randn(B, T, device=dev) * 0.063
out = torch.empty_like(x)
f = torch.fft.rfftfreq(NFFT, d=1 / SR).to(dev)
fc = torch.tensor([300.0, 600.0, 1000.0, 3000.0, 6000.0], device=dev)
A = torch.tensor([-10.0, -15.0, -20.0, -30.0, -40.0], device=dev)
ss = torch.zeros(B, dtype=torch.float64, device=dev)
sc = torch.ones(B, device=dev)
taps = bwe.get_FIR_lowpass(500, 1000, 1, SR).to(dev).reshape(-1)
cq = CQT_nsgt(7, 64, mode="oct", window=("kaiser", 1), fs=SR, audio_len=184184, device=dev)
BC = int(sys.argv[1]) if len(sys.argv) > 1 else 64
xc = torch.randn(BC, 184184, device=dev) * 0.063
fit = sampler.FilterFit(nfft=NFFT, sample_rate=SR, device=dev)
x8, y8 = x[:8].contiguous(), y[:8].contiguous()
abc = fit.stats(x8, y8)
p0 = torch.tensor([[280.0, 285, 290, 295, 300], [-15.0, -17, -20, -25, -30]], device=dev)
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
for _ in range(reps):
    ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, out=out)
    ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, adjoint=True, out=out)
    ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, sub=y, row_sumsq=ss, out=out)
    ops.apply_filter(x, NFFT, freqs=f, fc=fc, A=A, adjoint=True, row_scale=sc, out=out)
    ops.stft_stats(x, y, NFFT)
    ops.fir_filter(x, taps)
    fit(x8, y8, p0.clone(), abc=abc)
    cs = cq.fwd(xc.unsqueeze(1))
    cq.bwd(cs)
    cq.apply_hpf_DC(xc)
torch.cuda.synchronize()
print("ok")
