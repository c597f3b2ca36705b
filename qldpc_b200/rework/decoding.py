"""Drop-in for the reference's rework/decoding.py: normalised / damped / clipped min-sum, the 4-tuple
sum-product, the damped sum-product and OSD-w.  Same signatures, defaults and return conventions;
every decode is a batch of one in the CUDA kernels (csrc/bp_kernel.cuh, osd_kernel.cuh)."""
import numpy as np

from .._single import bp_single, osd_single
from ..code import cached_code


def performMinSum_Symmetric(H, syndrome, initialBelief, maxIter=50, alpha=1.0, damping=1.0, clip_llr=20.0,
                            alpha_estimation=False):
    """Reference: rework/decoding.py:5-75.
    Returns (candidateError int8[n], converged bool, posterior float64[n], currentIter int)."""
    if alpha_estimation:
        # :58-59 -- `return 0, 0, R_new / alpha, 0` right after the first check pass
        code = cached_code(H, "min_sum")
        R = code.bp_messages_batch(np.asarray(syndrome).reshape(1, -1), initialBelief, "min_sum", max(1, maxIter), alpha, damping,
                                   clip_llr if np.isfinite(clip_llr) else 1e300, dump_iter=0)[0]
        return 0, 0, R, 0
    return bp_single(H, syndrome, initialBelief, "min_sum", "min_sum", maxIter, alpha, damping, clip_llr)


def performBeliefPropagationFast(H, syndrome, initialBelief, maxIter=50):
    """Reference: rework/decoding.py:77-129 (sum-product, 4-tuple with the 0-based exit iteration)."""
    return bp_single(H, syndrome, initialBelief, "sum_product", "sum_product", maxIter)


def performBeliefPropagation_Symmetric(H, syndrome, initialBelief, maxIter=50, alpha=1.0, damping=0.8, clip_llr=20.0,
                                       alpha_estimation=False):
    """Reference: rework/decoding.py:131-191 (sum-product with alpha scaling, damping on Q, symmetric clip)."""
    if alpha_estimation:
        # :168-169 -- `return 0, 0, R, 0` at currentIter == 10 (and no early exit before, :188)
        if maxIter <= 10:
            raise ValueError("alpha_estimation=True returns the messages of iteration 10: maxIter must exceed 10")
        code = cached_code(H, "sum_product_sym")
        R = code.bp_messages_batch(np.asarray(syndrome).reshape(1, -1), initialBelief, "sum_product_sym", maxIter, alpha, damping,
                                   clip_llr if np.isfinite(clip_llr) else 1e300, dump_iter=10)[0]
        return 0, 0, R, 0
    return bp_single(H, syndrome, initialBelief, "sum_product_sym", "sum_product_sym", maxIter, alpha, damping, clip_llr)


def performOSD_enhanced(H, syndrome, llr, hard, order=0, max_combinations=None):
    """Reference: rework/decoding.py:193-278 (verbatim copy of decoding/OSD_enhanced.py:5-131)."""
    return osd_single(H, syndrome, llr, hard, order=order, max_combinations=max_combinations)
