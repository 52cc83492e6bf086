"""Oracle: octave-rasterised NSGT constant-Q transform (CPU, torch, fp64-capable).

TEST INFRASTRUCTURE -- see ``oracle/__init__.py``.

PARITY UNPINNED.  The reference imports ``CQT_nsgt`` from the third-party
package ``cqt_nsgt_pytorch`` (networks/cqtdiff+.py:9 of eloimoliner/BABE; PyPI
``cqt-nsgt-pytorch``, upstream eloimoliner/CQT_pytorch, NO version pin -- the
reference has no requirements file).  That package is absent from the
reference tree, not installed and not downloadable here, and the reference
holds no test or golden vector for it.  This file therefore restates the
PUBLISHED algorithm -- the painless non-stationary Gabor frame of Holighaus,
Doerfler, Velasco, Grill, "A framework for invertible, real-time constant-Q
transforms" (IEEE TASLP 2013), in the non-sliced form of Grill's ``nsgt``
package from which upstream descends -- and is anchored on

* the reference's call sites: constructor arguments networks/cqtdiff+.py:620,
  ``fwd`` :743, ``bwd`` :841-843, ``apply_hpf_DC``
  testing/blind_bwe_sampler.py:156, diff_params/edm.py:197;
* the structural contract those call sites impose: ``numocts`` tensors
  (B,1,binsoct,T_o) complex, lowest octave first, T_{o+1} = 2 T_o
  (networks/cqtdiff+.py:750,767-774), output length >= audio_len (:843);
* the transform's own invariants (tests/test_nsgt_oracle.py): partition of
  unity, perfect reconstruction up to the DC/Nyquist high-pass, adjointness.

Specification (every choice that upstream could have made differently is
marked [choice]):

 1. fmax = fs/2 - 1e-6, fmin = fmax / 2**numocts, K = numocts*binsoct bands at
    f_j = fmin * 2**(j*odiv), odiv = numocts/(K-1)  (Grill's LogScale: both end
    points included) [choice].  Bin positions fb_j = f_j * Ls / fs.
 2. Centre bins p_j = round(fb_j) for j < K-1; the last band sits half way to
    Nyquist, p_{K-1} = round((fb_{K-2} + Ls/2)/2)  (Grill, non-sliced).
 3. Window lengths Lg_j = max(4, round(fb_j / q)), q = 2**(odiv/2)/(2**odiv-1)/2
    (= distance between the neighbouring centres to first order) for j < K-1,
    Lg_{K-1} = max(4, round(Ls/2 - fb_{K-2})) [choice].  DC band: centre 0,
    Lg = max(4, 2 p_0); Nyquist band: centre Ls/2, Lg = max(4, 2 (Ls/2 - p_{K-1})).
 4. Window g_j[i], i < Lg_j, sits on bins p_j - Lg_j//2 + i; ("kaiser", beta) is
    the periodic Kaiser window I0(beta sqrt(1-(2i/Lg-1)^2))/I0(beta), "hann" the
    periodic Hann window.  Samples falling outside [0, Ls/2] are discarded
    [choice] (at most a rounding bin in the shipped configurations).
 5. "oct" mode: M_o = nextpow2(max_j-in-octave Lg_j); widened if needed so that
    M_{o+1} = 2 M_o.  DC / Nyquist bands keep M = Lg and are dropped from the
    output.
 6. fwd: X = FFT_Ls(x); c_j = IFFT_{M_o}(fold(X[p_j - Lg//2 + i] g_j[i])), fold
    index (i - Lg//2) mod M_o.
 7. duals gd_j = g_j / D, D[k] = sum over ALL bands (DC and Nyquist included)
    of M_j g_j[k]^2.
 8. bwd: FR[k] = sum_j M_o gd_j[i] FFT_{M_o}(c_j)[fold index]; x = irfft(FR, Ls).
 9. apply_hpf_DC(x) = irfft(rfft(x) * Hhpf), Hhpf[k] = sum_{CQ bands} M g gd,
    which equals 1 - (DC and Nyquist share).
"""
import math

import torch


def _nextpow2(v):
    return 1 << max(0, int(math.ceil(math.log2(max(1, v)))))


def _window(kind, L):
    i = torch.arange(L, dtype=torch.float64)
    if isinstance(kind, (tuple, list)) and kind[0] == "kaiser":
        beta = float(kind[1])
        r = 2.0 * i / L - 1.0
        return torch.special.i0(beta * torch.sqrt(torch.clamp(1.0 - r * r, min=0.0))) / \
            torch.special.i0(torch.tensor(beta, dtype=torch.float64))
    if kind == "hann":
        return 0.5 - 0.5 * torch.cos(2.0 * math.pi * i / L)
    if kind == "hamming":
        return 0.54 - 0.46 * torch.cos(2.0 * math.pi * i / L)
    raise NotImplementedError(f"window {kind!r}")


class NSGT:
    def __init__(self, numocts, binsoct, fs, Ls, window=("kaiser", 1), dtype=torch.float64):
        assert Ls % 2 == 0, "even signal lengths only"
        self.numocts, self.binsoct, self.fs, self.Ls, self.dtype = numocts, binsoct, fs, Ls, dtype
        K = numocts * binsoct
        Nc = Ls // 2
        fmax = fs / 2 - 1e-6
        fmin = fmax / 2 ** numocts
        odiv = numocts / (K - 1)
        q = 2 ** (odiv / 2) / (2 ** odiv - 1) / 2
        fb = [fmin * 2 ** (j * odiv) * Ls / fs for j in range(K)]
        p, Lg = [], []
        for j in range(K - 1):
            p.append(int(round(fb[j])))
            Lg.append(max(4, int(round(fb[j] / q))))
        p.append(int(round((fb[K - 2] + Nc) / 2)))
        Lg.append(max(4, int(round(Nc - fb[K - 2]))))
        # octave sizes
        M = []
        for o in range(numocts):
            M.append(_nextpow2(max(Lg[o * binsoct:(o + 1) * binsoct])))
        for o in range(numocts - 2, -1, -1):           # enforce doubling
            M[o] = max(M[o], M[o + 1] // 2)
        for o in range(1, numocts):
            M[o] = max(M[o], 2 * M[o - 1])
        self.p, self.Lg, self.M = p, Lg, M
        self.g = [_window(window, L) for L in Lg]
        # DC / Nyquist bands
        self.Lg_dc = max(4, 2 * p[0])
        self.Lg_ny = max(4, 2 * (Nc - p[K - 1]))
        g_dc, g_ny = _window(window, self.Lg_dc), _window(window, self.Lg_ny)
        # frame-operator diagonal on bins [0, Nc]
        D = torch.zeros(Nc + 1, dtype=torch.float64)

        def accumulate(centre, g, Mj):
            L = g.numel()
            k = centre - L // 2 + torch.arange(L)
            ok = (k >= 0) & (k <= Nc)
            D.index_add_(0, k[ok], Mj * g[ok] ** 2)

        accumulate(0, g_dc, self.Lg_dc)
        accumulate(Nc, g_ny, self.Lg_ny)
        for j in range(K):
            accumulate(p[j], self.g[j], M[j // binsoct])
        self.D = D
        self.gd = []
        hp = torch.zeros(Nc + 1, dtype=torch.float64)
        for j in range(K):
            L = Lg[j]
            k = p[j] - L // 2 + torch.arange(L)
            ok = (k >= 0) & (k <= Nc)
            gd = torch.zeros(L, dtype=torch.float64)
            gd[ok] = self.g[j][ok] / D[k[ok]]
            self.gd.append(gd)
            hp.index_add_(0, k[ok], M[j // binsoct] * self.g[j][ok] * gd[ok])
        self.Hhpf = hp

    # ------------------------------------------------------------------
    def _band_bins(self, j):
        L = self.Lg[j]
        k = self.p[j] - L // 2 + torch.arange(L)
        ok = (k >= 0) & (k <= self.Ls // 2)
        fold = (torch.arange(L) - L // 2) % self.M[j // self.binsoct]
        return k, ok, fold

    def fwd_loop(self, x):
        """Band-by-band statement of item 6 (slow; kept as the readable
        definition and checked against ``fwd`` in the tests)."""
        X = torch.fft.rfft(x.to(self.dtype), dim=-1)
        cdt = X.dtype
        out = []
        for o in range(self.numocts):
            Mo = self.M[o]
            buf = torch.zeros(x.shape[0], self.binsoct, Mo, dtype=cdt)
            for b in range(self.binsoct):
                j = o * self.binsoct + b
                k, ok, fold = self._band_bins(j)
                buf[:, b, fold[ok]] = X[:, k[ok]] * self.g[j][ok].to(self.dtype)
            out.append(torch.fft.ifft(buf, dim=-1))
        return out

    def bwd_loop(self, cs):
        """Band-by-band statement of item 8."""
        B = cs[0].shape[0]
        Nc = self.Ls // 2
        FR = torch.zeros(B, Nc + 1, dtype=cs[0].dtype)
        for o in range(self.numocts):
            Mo = self.M[o]
            C = torch.fft.fft(cs[o], dim=-1)
            for b in range(self.binsoct):
                j = o * self.binsoct + b
                k, ok, fold = self._band_bins(j)
                FR[:, k[ok]] += C[:, b, fold[ok]] * (Mo * self.gd[j][ok]).to(self.dtype)
        return torch.fft.irfft(FR, n=self.Ls, dim=-1)

    def _octave_tables(self, o):
        """Padded gather/scatter tables of one octave (same arithmetic as the
        loops above, vectorised over the bands so that the CPU baseline is not
        dominated by Python indexing)."""
        if not hasattr(self, "_tab"):
            self._tab = {}
        if o not in self._tab:
            Nc, Mo, nb = self.Ls // 2, self.M[o], self.binsoct
            Lmax = max(self.Lg[o * nb:(o + 1) * nb])
            i = torch.arange(Lmax)
            kk = torch.zeros(nb, Lmax, dtype=torch.long)
            fold = torch.zeros(nb, Lmax, dtype=torch.long)
            wg = torch.zeros(nb, Lmax, dtype=torch.float64)
            wd = torch.zeros(nb, Lmax, dtype=torch.float64)
            for b in range(nb):
                j = o * nb + b
                L = self.Lg[j]
                k = self.p[j] - L // 2 + i
                ok = (i < L) & (k >= 0) & (k <= Nc)
                kk[b] = k.clamp(0, Nc)
                fold[b] = (i - L // 2) % Mo            # distinct for i < Lmax <= Mo
                wg[b, :L] = self.g[j]
                wd[b, :L] = Mo * self.gd[j]
                wg[b] *= ok
                wd[b] *= ok
            self._tab[o] = (kk, fold, wg.to(self.dtype), wd.to(self.dtype))
        return self._tab[o]

    def fwd(self, x):
        """x (B, Ls) real -> list of numocts complex tensors (B, binsoct, M_o)."""
        X = torch.fft.rfft(x.to(self.dtype), dim=-1)
        B = x.shape[0]
        out = []
        for o in range(self.numocts):
            kk, fold, wg, _ = self._octave_tables(o)
            vals = X[:, kk] * wg                                        # (B, bands, Lmax)
            buf = torch.zeros(B, self.binsoct, self.M[o], dtype=X.dtype)
            buf = buf.scatter(2, fold[None].expand(B, -1, -1), vals)
            out.append(torch.fft.ifft(buf, dim=-1))
        return out

    def bwd(self, cs):
        """list of (B, binsoct, M_o) complex -> (B, Ls) real."""
        B = cs[0].shape[0]
        Nc = self.Ls // 2
        FR = torch.zeros(B, Nc + 1, dtype=cs[0].dtype)
        for o in range(self.numocts):
            kk, fold, _, wd = self._octave_tables(o)
            C = torch.fft.fft(cs[o], dim=-1)
            vals = C.gather(2, fold[None].expand(B, -1, -1)) * wd
            FR = FR.index_add(1, kk.reshape(-1), vals.reshape(B, -1))
        return torch.fft.irfft(FR, n=self.Ls, dim=-1)

    def apply_hpf_DC(self, x):
        X = torch.fft.rfft(x.to(self.dtype), dim=-1)
        return torch.fft.irfft(X * self.Hhpf.to(self.dtype), n=self.Ls, dim=-1)
