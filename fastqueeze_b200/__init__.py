"""fastqueeze_b200 -- B200-native implementation of the SeqArc v1.6 entropy-coding hot path.

Only what the path needs lives here: `csrc/` (hand-written sm_100a CUDA kernels + the C ABI of
include/seqarc_b200.h) and `codec.py` (host-side mirror of the reference's EncapFqzComp seam).
"""
from ._lib import LIB_PATH, SeqArcB200Error, lib  # noqa: F401
from .codec import DebugInfo, EncapFqzComp, SeqArcMemBuf, SeqArcParam  # noqa: F401
