"""Drop-in for the third-party ``cqt_nsgt_pytorch`` package the reference
imports (``from cqt_nsgt_pytorch import CQT_nsgt``, networks/cqtdiff+.py:9 of
eloimoliner/BABE).  The class lives in ``babe_b200.cqt`` and runs on the
sm_100a kernels; this package only provides the import name.

PARITY UNPINNED versus upstream (the package is not available offline): the transform follows
the specification in oracle/nsgt.py.  If the real upstream distribution is installed in the environment this
package SHADOWS it whenever the repository root precedes site-packages on ``sys.path`` -- a warning says so, because
weights trained with upstream's band layout must be validated against this transform before use."""
import warnings

from babe_b200.cqt import CQT_nsgt  # noqa: F401

__all__ = ["CQT_nsgt"]


def _warn_if_shadowing():
    try:
        from importlib import metadata
        for name in ("cqt-nsgt-pytorch", "cqt_nsgt_pytorch"):
            try:
                v = metadata.version(name)
            except metadata.PackageNotFoundError:
                continue
            warnings.warn(f"babe_b200's cqt_nsgt_pytorch shadows the installed upstream distribution {name} {v}: "
                          "CQT_nsgt here follows the in-repo NSGT specification (oracle/nsgt.py), whose parity with "
                          "upstream is unpinned; validate pretrained CQTDiff+ weights before relying on it",
                          RuntimeWarning, stacklevel=3)
            return
    except Exception:                                   # noqa: BLE001 - never fail an import over a courtesy check
        pass


_warn_if_shadowing()
