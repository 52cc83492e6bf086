"""babe_b200 -- B200-native signal-processing hot path of BABE blind bandwidth
extension (eloimoliner/BABE), behind the reference's own Python seams.

* ``babe_b200.blind_bwe_utils``  drop-in for ``utils.blind_bwe_utils``
* ``cqt_nsgt_pytorch.CQT_nsgt``  (top-level package in this repo) drop-in for
  the third-party CQT the reference imports at networks/cqtdiff+.py:9
* ``babe_b200.sampler.BlindSamplerFused``  drop-in ``tester.sampler_callable``
* ``babe_b200.install()``        route the reference's imports to the above

All arithmetic runs in hand-written sm_100a CUDA kernels loaded through the C
ABI of ``include/babe_b200.h`` (``libbabe_b200.so``); there is no CPU fallback.
"""
import sys

__version__ = "0.1.0"


def install():
    """Make ``import utils.blind_bwe_utils`` (as done by
    testing/blind_bwe_sampler.py:9 and testing/blind_bwe_tester.py:26) resolve
    to the CUDA drop-in.  ``cqt_nsgt_pytorch`` needs no patching: the package of
    that name at the repository root is found by a normal import."""
    from . import blind_bwe_utils
    sys.modules["utils.blind_bwe_utils"] = blind_bwe_utils
    pkg = sys.modules.get("utils")
    if pkg is not None:
        setattr(pkg, "blind_bwe_utils", blind_bwe_utils)
    return blind_bwe_utils
