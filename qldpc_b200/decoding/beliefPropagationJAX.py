"""Names of the reference's decoding/beliefPropagationJAX.py (float32 / int32 conventions), served
by the same CUDA kernels (float32 instantiation).  The reference module itself is broken for
non-square H (:60 multiplies (n, m) by (n,)); the semantics implemented here are those of
performBeliefPropagationFast, which that module set out to restate."""
import numpy as np

from .._single import bp_single
from ..code import cached_code
from .beliefPropagationGPU import generate_errors_and_syndromes_batch


def performBeliefPropagationJAX(H, syndrome, initialBelief, verbose=False, maxIter=50):
    """Reference: beliefPropagationJAX.py:107-119 -> (int8[n], bool, float32[n])."""
    hard, ok, llr, _ = bp_single(H, syndrome, initialBelief, "sum_product", "sum_product", maxIter, precision=32)
    return hard, ok, llr.astype(np.float32)


def performBeliefPropagationBatchJAX(H, syndromes, initialBelief, maxIter=50):
    """Reference: beliefPropagationJAX.py:122-145 -> (int8 (B, n), bool (B,), float32 (B, n))."""
    code = cached_code(H, "sum_product")
    hard, conv, llr, _ = code.bp_decode_batch(np.asarray(syndromes), initialBelief, variant="sum_product", max_iter=maxIter,
                                              precision=32, want_llr=True)
    return hard, conv, llr.astype(np.float32)


def generate_errors_and_syndromes(H, error_rate, batch_size, key=None):
    """Reference: beliefPropagationJAX.py:148-157 (JAX PRNG key -> here an int seed or None).
    Returns (errors int32 (B, n), syndromes int32 (B, m))."""
    rng = None if key is None else np.random.default_rng(int(key))
    e, s = generate_errors_and_syndromes_batch(H, error_rate, batch_size, rng)
    return e.astype(np.int32), s.astype(np.int32)


def warmup_jit(H, batch_size=100, maxIter=50):
    """Reference: beliefPropagationJAX.py:161-174.  Nothing is traced here; this builds the device
    tables of H and runs one small batch so that later calls are steady-state."""
    code = cached_code(H, "sum_product")
    code.bp_decode_batch(np.zeros((min(batch_size, 32), code.m), np.uint8), np.full(code.n, 3.0), variant="sum_product",
                         max_iter=min(maxIter, 2), precision=32, want_llr=False)
