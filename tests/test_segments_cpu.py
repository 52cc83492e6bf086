"""CPU: host-side segmentation / cross-fade logic for long recordings
(testing/blind_bwe_tester.py:413-577, non-AR branch)."""
import torch

from conftest import rel_l2

from babe_b200 import segments


def test_spans_match_reference_loop():
    # transliteration of the index arithmetic at testing/blind_bwe_tester.py:421-521
    L, segL, OLA, discard_end, discard_start = 1323000, 184184, 256, 200, 0
    ix, ref = 0, [0]
    ix += segL - discard_end - OLA
    while ix < L - segL - discard_end - discard_start:
        ref.append(ix)
        ix += segL - discard_end - OLA
    ref.append(ix)
    spans = segments.segment_spans(L, segL, OLA, discard_end)
    assert [a for a, _ in spans] == ref
    assert len(spans) == 8                    # BASELINE config 5: 60 s -> 8 segments, one per GPU


def test_identity_sampler_reconstructs():
    torch.manual_seed(0)
    L, segL = 50000, 8192
    x = torch.randn(1, L)
    segs, spans = segments.split(x, segL)
    assert segs.shape == (len(spans), segL)
    y = segments.merge(segs, spans, L)
    # hann(2*OLA)[:OLA] + hann(2*OLA)[OLA:] == 1: the cross-fade is transparent
    assert torch.allclose(y, x, atol=1e-6)


class _FakeSampler:
    joint = False

    def predict_blind_bwe(self, y, rid=False):
        return 2.0 * y, torch.ones(2, 5) * y.shape[0]


def test_restore_recording_single_process():
    x = torch.randn(1, 30000)
    out, filt = segments.restore_recording(_FakeSampler(), x, 4096)
    assert torch.allclose(out, 2.0 * x, atol=1e-5)
    assert len(filt) == len(segments.segment_spans(30000, 4096)) and filt[0][1].shape == (2, 5)
    out_j, filt_j = segments.restore_recording(_FakeSampler(), x, 4096, joint=True)
    assert torch.allclose(out_j, 2.0 * x, atol=1e-5)


class _StubSampler:
    """Records the calls of the AR driver; 'restoration' = 2 x the lowpassed input outside the mask."""

    def __init__(self):
        self.calls = []

    def predict_blind_bwe(self, y, rid=False):
        self.calls.append(("blind", tuple(y.shape)))
        return y.clone(), torch.tensor([[500.0], [-20.0]])

    def predict_bwe(self, seg, filt, typefilter, rid=False):
        self.calls.append(("bwe", tuple(seg.shape)))
        return 2 * seg

    def predict_bwe_AR(self, seg, y_masked, filt, typefilter, rid=False, mask=None):
        self.calls.append(("ar", int(mask.sum())))
        return mask * y_masked + (1 - mask) * 2 * seg


def test_restore_recording_ar_driver():
    """testing/blind_bwe_tester.py:710-867: window placement, known-head masks, last padded window."""
    import numpy as np
    from babe_b200 import segments
    sr, segL, ov_s = 1000, 400, 0.05                     # overlap 50 samples, discard_end 200 -> hop 150
    L = 1730
    g = torch.Generator().manual_seed(3)
    degraded = torch.randn(1, L, generator=g)
    s = _StubSampler()
    out, filt = segments.restore_recording_ar(s, degraded.clone(), segL, sr, overlap_s=ov_s,
                                              n_segments_blindstep=2, std=0.1,
                                              rng=np.random.RandomState(0))
    assert out.shape == degraded.shape and filt.shape == (2, 1)
    kinds = [c[0] for c in s.calls]
    # AR windows start at 150, 300, ... while ix < L - segL - 200 = 1130 (7 of them) or more than one window
    # is left (1200: 530 samples left), then the last, zero padded one at 1350
    assert kinds[0] == "blind" and s.calls[0][1] == (2, segL)
    assert kinds[1] == "bwe" and kinds[2:] == ["ar"] * 9
    assert all(c[1] == 50 for c in s.calls[2:])           # 50 known samples at the head of every window
    # the stub doubles every unknown sample and copies the known head: the result is 2 x the input (the
    # scale normalisation is undone at the end) -- except the head of the LAST window, which the
    # reference fills with the tail of the previous full prediction (:842, kept as is)
    ok = torch.ones(L, dtype=torch.bool)
    ok[1350:1400] = False
    assert rel_l2(out[:, ok], 2 * degraded[:, ok]) < 1e-5
    assert rel_l2(out[:, 1350:1400], 2 * degraded[:, 1550:1600]) < 1e-5
    # a given filter skips the blind step
    s2 = _StubSampler()
    segments.restore_recording_ar(s2, degraded.clone(), segL, sr, overlap_s=ov_s,
                                  estimated_filter=torch.tensor([[500.0], [-20.0]]))
    assert [c[0] for c in s2.calls][0] == "bwe"
