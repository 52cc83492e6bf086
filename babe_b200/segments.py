"""Long recordings as overlapped fixed-length segments (host orchestration).

Restates the non-autoregressive branch of ``BlindTester.formal_test_bwe``
(testing/blind_bwe_tester.py:413-577 of eloimoliner/BABE, ``use_AR: False``,
``OLA: 256`` in the formal configs) without soundfile / wandb: the recording is
cut into ``audio_len`` windows that advance by ``segL - discard_end - OLA``
samples, every window is restored by an independent sampler call (its own
filter estimate, :433,:477,:543) and the predictions are cross-faded with a
``2*OLA`` Hann window.  Segments are independent units, so they shard over
ranks with no data-path collective; ``restore_recording`` gathers
predictions and per-segment filters with NCCL at the end.
"""
import torch

from . import distributed as bd


def segment_spans(L, seg_len, ola=256, discard_end=200, discard_start=0):
    """Start indices exactly as the while-loop at testing/blind_bwe_tester.py:469-521
    produces them; the last span is the incomplete tail (:526-533)."""
    step = seg_len - discard_end - ola
    spans, ix = [(0, seg_len)], step
    while ix < L - seg_len - discard_end - discard_start:
        spans.append((ix, ix + seg_len))
        ix += step
    spans.append((ix, ix + seg_len))          # tail, zero padded to seg_len
    return spans


def split(degraded, seg_len, ola=256, discard_end=200):
    """degraded (1, L) -> (segments (S, seg_len), spans)."""
    L = degraded.shape[-1]
    spans = segment_spans(L, seg_len, ola, discard_end)
    segs = degraded.new_zeros((len(spans), seg_len))
    for s, (a, b) in enumerate(spans):
        piece = degraded[0, a:min(b, L)]
        segs[s, :piece.shape[-1]] = piece
    return segs, spans


def merge(preds, spans, L, ola=256, discard_end=200):
    """Cross-fade of the per-segment predictions (:455-461, :493-499, :565-568)."""
    seg_len = preds.shape[-1]
    hann = torch.hann_window(ola * 2, device=preds.device)
    out = preds.new_zeros((1, L))
    last = len(spans) - 1
    for s, (a, _) in enumerate(spans):
        if s == last:
            n = L - a
            w = preds[s, :n].clone()
            w[:ola] *= hann[:ola]
            out[0, a:] += w
            continue
        w = preds[s, :seg_len - discard_end].clone()
        if s > 0:
            w[:ola] *= hann[:ola]
        w[-ola:] *= hann[ola:]
        out[0, a:a + seg_len - discard_end] += w
    return out


def restore_recording(sampler, degraded, seg_len, ola=256, discard_end=200, joint=False):
    """Blind restoration of a whole recording, segments sharded over the ranks of the
    current process group.  Returns (final_pred (1, L), filter_data) on every rank,
    ``filter_data`` being the list [((ix0, ix1), filter_params), ...] the reference
    pickles (:466-467,:576-577).  ``joint=True`` processes a rank's segments as one
    batch and shares ONE filter across all ranks (testing/blind_bwe_tester.py:758-781)."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    segs, spans = split(degraded, seg_len, ola, discard_end)
    lo, hi = bd.shard_rows(segs.shape[0], rank, world)
    mine = segs[lo:hi]
    if joint:
        sampler.joint = True
        pred, filt = sampler.predict_blind_bwe(mine, rid=False)
        filts = filt.unsqueeze(0).expand(mine.shape[0], -1, -1).contiguous()
    else:
        preds, fl = [], []
        for s in range(mine.shape[0]):
            p, f = sampler.predict_blind_bwe(mine[s:s + 1].clone(), rid=False)
            preds.append(p)
            fl.append(f)
        pred = torch.cat(preds, 0) if preds else mine.new_zeros((0, seg_len))
        filts = torch.stack(fl, 0) if fl else mine.new_zeros((0, 2, 1))
    all_pred = bd.gather_rows(pred)
    all_filt = bd.gather_rows(filts)
    final = merge(all_pred, spans, degraded.shape[-1], ola, discard_end)
    return final, [(spans[s], all_filt[s]) for s in range(len(spans))]
