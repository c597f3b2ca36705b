"""qldpc_b200 -- B200-native (sm_100a) BP+OSD decoder for quantum LDPC codes.

Drop-in for the decode hot path of michelebanfi/qLDPC: the same function names and return
conventions (qldpc_b200.decoding.*, qldpc_b200.rework.decoding, qldpc_b200.spaceTime) on top of a
batched C-ABI CUDA library (include/qldpc_b200.h).  There is no CPU fallback.
"""
from ._lib import QldpcError, BPConfig, COUNTER_NAMES  # noqa: F401
from .code import Code, load_code, cached_code  # noqa: F401

# Arithmetic of the reference-named single-shot wrappers: 64 = the reference's float64 (bit-exact
# min-sum), 32 = the production float32 kernels.  The batched Code.* methods take `precision=`.
DEFAULT_PRECISION = 64
