"""Drop-in for the reference's decoding/beliefPropagationGPU.py (CuPy batch BP) -- same names,
hand-written sm_100a kernels instead of dense (B, m, n) float64 CuPy ufuncs, and no host round
trip per iteration (reference :157-167)."""
import numpy as np

import qldpc_b200 as _pkg
from .._single import bp_single
from ..code import cached_code, to_dense

GPU_AVAILABLE = True   # reference :11 -- here the GPU is mandatory, there is no NumPy fallback


def performBeliefPropagationGPU(H, syndrome, initialBelief, verbose=False, maxIter=50):
    """Reference: beliefPropagationGPU.py:22-78.  (int8[n], bool, float64[n])."""
    hard, ok, llr, it = bp_single(H, syndrome, initialBelief, "sum_product", "sum_product", maxIter)
    if verbose and ok:
        print(f"Error found at iteration {it}")
    return hard, ok, llr


def performBeliefPropagationBatch(H, syndromes, initialBelief, maxIter=50):
    """Reference: beliefPropagationGPU.py:81-178.
    syndromes (B, m) -> (candidateErrors int8 (B, n), converged bool (B,), values float64 (B, n));
    each shot is frozen at its first convergence (:160-167), non-converged shots return the last iterate."""
    code = cached_code(H, "sum_product")
    hard, conv, llr, _ = code.bp_decode_batch(np.asarray(syndromes), initialBelief, variant="sum_product",
                                              max_iter=maxIter, precision=_pkg.DEFAULT_PRECISION, want_llr=True)
    return hard, conv, llr


def generate_errors_and_syndromes_batch(H, error_rate, batch_size, rng=None):
    """Reference: beliefPropagationGPU.py:181-200.  With an explicit NumPy Generator the draws are the
    reference's (`rng.random((B, n)) < p`) so seeded scripts reproduce; the syndromes are then computed
    from those errors.  With rng=None the errors are sampled on the device (Philox4x32-10).
    Returns (errors int8 (B, n), syndromes int8 (B, m))."""
    Hd = to_dense(H)
    if rng is None:
        code = cached_code(Hd, "sum_product")
        seed = int(np.random.SeedSequence().generate_state(2, np.uint32).view(np.uint64)[0])
        return code.sample(error_rate, batch_size, seed=seed)
    num_checks, num_vars = Hd.shape
    errors = (rng.random((batch_size, num_vars)) < error_rate).astype(np.int8)       # the reference's draws (:195)
    return errors, cached_code(Hd, "sum_product").syndromes(errors)                  # err * H^T mod 2 on the device (:198)
