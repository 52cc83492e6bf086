"""Times the spectrogram-distance kernels of the STFT-guidance branches (a9 / a10) at the bench shape
(B = 8 and 64 chains x T = 184184, NFFT 4096 -> F = 2049, 88 frames), CUDA events, L2 flushed between iterations,
against the algorithmic bytes (stats: both spectrograms read once; grad: both read, one gradient written)."""
import json
import sys
sys.path.insert(0, ".")
import torch
from babe_b200 import ops

PEAK = json.load(open("MEASURED_PEAKS.json")).get("hbm_GBps", 6550.4) if False else 6550.4
dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


for B in (8, 64):
    F, M = 2049, 88
    X = torch.randn(B, F, M, 2, device=dev)
    R = torch.randn(B, F, M, 2, device=dev)
    w = torch.linspace(0, 1, F, device=dev)
    coef = torch.ones(1, device=dev)
    nb = X.numel() * 4
    for name, fn, byt in (
            ("spec_dist_stats mode0", lambda: ops.spec_dist_stats(X, R, w, 0), 2 * nb),
            ("spec_dist_stats mode2", lambda: ops.spec_dist_stats(X, R, w, 2), 2 * nb),
            ("spec_mag_stats", lambda: ops.spec_mag_stats(X, R, None, w), 2 * nb),
            ("spec_dist_grad mode0", lambda: ops.spec_dist_grad(X, R, w, coef, 0), 3 * nb),
            ("spec_dist_grad mode2", lambda: ops.spec_dist_grad(X, R, w, coef, 2), 3 * nb),
            ("spec_mag_grad", lambda: ops.spec_mag_grad(X, R, None, w, coef), 3 * nb)):
        ms = timed(fn)
        print(json.dumps({"op": name, "B": B, "F": F, "frames": M, "ms": round(ms, 4), "MB": round(byt / 1e6, 1),
                          "GBps": round(byt / ms / 1e6, 1), "frac_hbm": round(byt / ms / 1e6 / PEAK, 3)}))
