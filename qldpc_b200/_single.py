"""Shared plumbing of the reference-named single-shot wrappers (B = 1 batches on the GPU)."""
import numpy as np

import qldpc_b200 as _pkg
from .code import cached_code, _bits


def bp_single(H, syndrome, initialBelief, family, variant, maxIter, alpha=1.0, damping=1.0, clip=20.0, precision=None,
              schedule=None):
    code = cached_code(H, family, schedule)
    synd = _bits(np.asarray(syndrome).reshape(1, -1))
    hard, conv, llr, iters = code.bp_decode_batch(synd, initialBelief, variant=variant, max_iter=maxIter, alpha=alpha,
                                                  damping=damping, clip=clip,
                                                  precision=precision or _pkg.DEFAULT_PRECISION, want_llr=True)
    return hard[0], bool(conv[0]), llr[0], int(iters[0])


def osd_single(H, syndrome, llr, hard, order=0, max_combinations=None):
    code = cached_code(H, "loop")
    out = code.osd_decode_batch(np.asarray(syndrome).reshape(1, -1), np.asarray(llr, dtype=np.float64).reshape(1, -1),
                                np.asarray(hard).reshape(1, -1), order=order, max_combinations=max_combinations)
    return out[0]
