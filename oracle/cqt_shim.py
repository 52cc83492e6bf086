"""CPU stand-in with the ``CQT_nsgt`` interface, backed by the oracle NSGT.
TEST INFRASTRUCTURE (see oracle/__init__.py): lets the reference network file and the in-repo denoiser
restatement run on the CPU against the same transform."""
import torch

from .nsgt import NSGT


class OracleCQT:
    def __init__(self, numocts, binsoct, mode="oct", window=("kaiser", 1), flex_Q=None, fs=44100,
                 audio_len=44100, device="cpu", dtype=torch.float32):
        assert mode == "oct"
        self.t = NSGT(numocts, binsoct, fs, audio_len, window, dtype=dtype)
        self.Ls = audio_len

    def fwd(self, x):                       # (B,1,T) -> list of (B,1,bins,T_o)
        return [c.unsqueeze(1) for c in self.t.fwd(x.squeeze(1))]

    def bwd(self, cs):                      # -> (B,1,T)
        return self.t.bwd([c.squeeze(1) for c in cs]).unsqueeze(1)

    def apply_hpf_DC(self, x):
        return self.t.apply_hpf_DC(x)
