"""Drop-in for the reference's decoding/beliefPropagation.py (sum-product BP, single shot).

Same signatures, defaults and return conventions; the decode runs in the CUDA kernel
bp_decode_kernel<.., VAR_SUM_PRODUCT, ..> (qldpc_b200/csrc/bp_kernel.cuh) as a batch of one.
"""
import numpy as np

from .._single import bp_single


def performBeliefPropagation(H, syndrome, initialBelief, verbose=True, plotPath=None, maxIter=50):
    """Reference: decoding/beliefPropagation.py:6-85 (loop version; accepts scipy.sparse H).
    Returns (candidateError int8[n], converged bool, posterior LLR float64[n]).
    `plotPath` (Tanner-graph PNG, :30-31) is outside the decode path and is ignored."""
    if verbose:
        print(f"Initial syndrome: {np.asarray(syndrome, dtype=np.int8)}")
    hard, ok, llr, it = bp_single(H, syndrome, initialBelief, "loop", "sum_product", maxIter)
    if verbose and ok:
        print(f"Error found at iteration {it}: {hard}")
    return hard, ok, llr


def performBeliefPropagationFast(H, syndrome, initialBelief, verbose=True, maxIter=50):
    """Reference: decoding/beliefPropagation.py:88-144 (dense-vectorised version).
    Returns (candidateError int8[n], converged bool, posterior LLR float64[n])."""
    hard, ok, llr, it = bp_single(H, syndrome, initialBelief, "sum_product", "sum_product", maxIter)
    if verbose and ok:
        print(f"Error found at iteration {it}: {hard}")
    return hard, ok, llr
