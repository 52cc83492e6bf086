"""Drop-in for the third-party ``cqt_nsgt_pytorch`` package the reference
imports (``from cqt_nsgt_pytorch import CQT_nsgt``, networks/cqtdiff+.py:9 of
eloimoliner/BABE).  The class lives in ``babe_b200.cqt`` and runs on the
sm_100a kernels; this package only provides the import name."""
from babe_b200.cqt import CQT_nsgt  # noqa: F401

__all__ = ["CQT_nsgt"]
