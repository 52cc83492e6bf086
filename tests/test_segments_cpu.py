"""CPU: host-side segmentation / cross-fade logic for long recordings
(testing/blind_bwe_tester.py:413-577, non-AR branch)."""
import torch

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
