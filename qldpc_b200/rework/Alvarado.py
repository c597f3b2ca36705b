"""Mirror of the reference's rework/Alvarado.py:10-66 -- per-p estimation of the min-sum normalisation alpha
from the first-iteration check-to-variable messages, batched on the GPU."""
import numpy as np

from ..code import cached_code


def estimate_alpha_from_code(code, trials=5000, error_rate=0.05, maxIter=50, bins=50, verbose=True):
    """Reference: rework/Alvarado.py:10-66.  Draws `trials` errors from NumPy's global RNG exactly like the reference
    (one np.random.random(n) per trial), collects the messages R_new/alpha of performMinSum_Symmetric(alpha=1, damping=1,
    clip_llr=inf, alpha_estimation=True) split by the true bit, histograms them (density, common range), and fits
    log(hist_0 / hist_1) = alpha * lambda through the origin (scipy.optimize.curve_fit)."""
    from scipy.optimize import curve_fit
    H = np.asarray(code)
    n = H.shape[1]
    handle = cached_code(H, "min_sum")
    prior = np.full(n, np.log((1 - error_rate) / error_rate))
    errors = np.array([(np.random.random(n) < error_rate) for _ in range(trials)], dtype=np.uint8)
    synd = handle.syndromes(errors)
    edge_rows, edge_cols = np.nonzero(H)
    true_0, true_1 = [], []
    step = 4096
    for o in range(0, trials, step):
        R = handle.bp_messages_batch(synd[o:o + step], prior, "min_sum", max(1, maxIter), 1.0, 1.0, np.inf, dump_iter=0)
        msgs = R[:, edge_rows, edge_cols]                       # (b, E) in np.nonzero (row-major) order, as the reference
        bits = errors[o:o + step][:, edge_cols]
        true_0.append(msgs[bits == 0])
        true_1.append(msgs[bits == 1])
    true_0, true_1 = np.concatenate(true_0), np.concatenate(true_1)
    hist_range = (min(true_0.min(), true_1.min()), max(true_0.max(), true_1.max()))
    hist_0, bin_edges = np.histogram(true_0, bins=bins, range=hist_range, density=True)
    hist_1, _ = np.histogram(true_1, bins=bins, range=hist_range, density=True)
    centers = (bin_edges[:-1] + bin_edges[1:]) / 2
    ok = (hist_0 > 0) & (hist_1 > 0)
    popt, _ = curve_fit(lambda x, alpha: alpha * x, centers[ok], np.log(hist_0[ok] / hist_1[ok]))
    if verbose:
        print(f"Estimated alpha for error rate {error_rate}: {popt[0]}")
    return popt[0]
