"""Launch accounting and (optional) CUDA-event timing of the C-ABI calls.

Every wrapper in ``ops.py`` / ``cqt.py`` runs its library call inside
``op(name, kernels, algorithmic_bytes)``.  The launch counter is always on;
when ``enable()`` has been called each call is also bracketed by two CUDA
events on the launching stream, so that ``summary()`` can report per-operator
device time and achieved algorithmic bandwidth for the region between
``reset()`` and ``summary()`` (bench.py's ``roofline`` object).
"""
import contextlib

import torch

_enabled = False
_records = {}      # name -> [(start, end, bytes), ...]
_launches = 0


def enable(flag=True):
    global _enabled
    _enabled = flag


def reset():
    global _launches
    _records.clear()
    _launches = 0


def launches():
    return _launches


def add_launches(n):
    """Kernels replayed from a captured CUDA graph (the wrappers are not re-entered on replay)."""
    global _launches
    _launches += n


@contextlib.contextmanager
def op(name, kernels, nbytes):
    global _launches
    _launches += kernels
    if not _enabled:
        yield
        return
    s = torch.cuda.Event(enable_timing=True)
    e = torch.cuda.Event(enable_timing=True)
    s.record()
    yield
    e.record()
    _records.setdefault(name, []).append((s, e, nbytes, kernels))


def summary():
    """{name: dict(calls, kernels, ms_total, ms_avg, bytes_avg, gbs)}; call after a
    device synchronize."""
    out = {}
    for name, recs in _records.items():
        ms = [s.elapsed_time(e) for s, e, _, _ in recs]
        nbytes = [b for _, _, b, _ in recs]
        tot = sum(ms)
        out[name] = dict(calls=len(recs), kernels=sum(k for _, _, _, k in recs), ms_total=tot,
                         ms_avg=tot / len(recs), bytes_avg=sum(nbytes) / len(recs),
                         gbs=(sum(nbytes) / 1e9) / (tot / 1e3) if tot > 0 else 0.0)
    return out
