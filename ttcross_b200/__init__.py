"""ttcross_b200 — B200-native TT-cross sweep (dtt_dmrgg of aukeschaap/ttcross) behind a C-ABI.

The product is `libttcross_b200.so` (hand-written sm_100a CUDA + a C++ host engine, see csrc/ and
include/ttcross_b200.h).  This package is the thin Python host mirror used by tests and bench.py:
ctypes bindings (`api`) and the reference drivers' problem setup (`drivers`).  There is no CPU fallback.
"""
from .api import TTCross, TTCrossError, load_library, fp64_peak, qr_thin, tt_write, tt_read, ISING, STDNORM, MVN, COSCOEF  # noqa: F401
from . import drivers  # noqa: F401
from . import multi  # noqa: F401
