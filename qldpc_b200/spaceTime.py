"""Drop-in for the reference's spaceTime.py: the phenomenological space-time check matrix and its
syndrome sampler (inputs of BASELINE config 4).  Host-side construction of decoder INPUTS; the
decode of the resulting 864 x 2592 matrix runs in the HBM-staged BP kernel."""
import numpy as np


def spaceTimeMatrix(H, n_cycles):
    """Reference: spaceTime.py:4-18.  [ I_T (x) H | I_{mT} + (I shifted down by m) ], float64, C order."""
    H = np.asarray(H)
    m, n = H.shape
    T = int(n_cycles)
    out = np.zeros((m * T, n * T + m * T), dtype=np.float64)
    for t in range(T):
        out[t * m:(t + 1) * m, t * n:(t + 1) * n] = H
    diag = np.arange(m * T)
    out[diag, n * T + diag] = 1.0
    if T > 1:
        out[diag[m:], n * T + diag[:-m]] = 1.0      # np.eye(mT, k=-m): round t also sees the flip of round t-1
    return out


def spacetimeSyndrome(code, error_rate, n_cycles):
    """Reference: spaceTime.py:20-43, drawing from NumPy's global RNG in the same order (one length-n draw,
    then one length-m draw per cycle).  Measurement flips persist from round to round (:28-32); block 0 of
    the returned vector is the LAST round's syndrome (:35), blocks 1.. are consecutive differences."""
    code = np.asarray(code)
    m, n = code.shape
    error = (np.random.random(n) < error_rate).astype(int)
    syndrome = (error @ code.T) % 2
    history = []
    for _ in range(n_cycles):
        s_error = (np.random.random(m) < error_rate).astype(int)
        syndrome = (syndrome + s_error) % 2
        history.append(syndrome)
    blocks = [syndrome] + [(history[i] + history[i - 1]) % 2 for i in range(1, n_cycles)]
    return error, np.concatenate(blocks)
