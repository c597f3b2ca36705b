"""Mirror of the reference's rework/Alvarado.py:10-66 -- per-p estimation of the min-sum normalisation alpha from the
first-iteration check-to-variable messages -- computed on the GPU without materialising a single message."""
import numpy as np

from ..code import cached_code


def alpha_from_counts(counts, L, bins=50):
    """The tail of estimate_alpha_from_code (Alvarado.py:40-64) for messages that only take the two values +-L:
    counts[bit][s] = number of messages (1 - 2 s) L on edges whose variable has true value `bit`.  Reproduces
    np.histogram(density=True, range=(min, max)) and scipy's curve_fit of log(hist_0 / hist_1) = alpha * lambda."""
    from scipy.optimize import curve_fit
    counts = np.asarray(counts, dtype=np.float64).reshape(2, 2)
    vals = np.array([L, -L])                                   # s = 0 -> +L, s = 1 -> -L
    present = counts.sum(0) > 0
    lo, hi = vals[present].min(), vals[present].max()
    if lo == hi:                                               # np.histogram widens an empty range to (x - 0.5, x + 0.5)
        lo, hi = lo - 0.5, hi + 0.5
    edges = np.linspace(lo, hi, bins + 1)
    width = np.diff(edges)
    hist = np.zeros((2, bins))
    for s in range(2):
        if counts[:, s].sum() == 0:
            continue
        k = min(int(np.searchsorted(edges, vals[s], side="right")) - 1, bins - 1)       # the last bin is closed on the right
        hist[:, k] += counts[:, s]
    hist_0 = hist[0] / counts[0].sum() / width
    hist_1 = hist[1] / counts[1].sum() / width
    centers = (edges[:-1] + edges[1:]) / 2
    ok = (hist_0 > 0) & (hist_1 > 0)
    popt, _ = curve_fit(lambda x, alpha: alpha * x, centers[ok], np.log(hist_0[ok] / hist_1[ok]))
    return popt[0]


def estimate_alpha_from_code(code, trials=5000, error_rate=0.05, maxIter=50, bins=50, verbose=True, seed=None):
    """Reference: rework/Alvarado.py:10-66.  The reference collects R_new / alpha of performMinSum_Symmetric(alpha=1,
    damping=1, clip_llr=inf, alpha_estimation=True) -- the messages after the first check pass, whatever maxIter -- split by
    the true bit, histograms them (density, common range) and fits log(hist_0 / hist_1) = alpha * lambda through the origin.
    With the uniform prior L = ln((1-p)/p) every such message is (1 - 2 s_c) L, so the two histograms are four edge
    counts, accumulated on the device (Code.alpha_counts); the fit then runs on the two occupied bins exactly as the
    reference's does.  seed=None draws the errors from NumPy's global RNG exactly like the reference (one
    np.random.random(n) per trial) and reproduces its alpha to 1e-9; an integer seed samples them on the device."""
    H = np.asarray(code)
    n = H.shape[1]
    handle = cached_code(H, "min_sum")
    L = np.log((1 - error_rate) / error_rate)
    if seed is None:
        errors = np.array([(np.random.random(n) < error_rate) for _ in range(trials)], dtype=np.uint8)
        counts = handle.alpha_counts(errors)
    else:
        counts = handle.alpha_counts(None, p=error_rate, nshots=trials, seed=seed)
    alpha = alpha_from_counts(counts, L, bins)
    if verbose:
        print(f"Estimated alpha for error rate {error_rate}: {alpha}")
    return alpha
