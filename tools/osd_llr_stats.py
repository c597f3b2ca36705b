"""How many |LLR| of a BP-failed shot on the 864 x 2592 space-time matrix are below the shot's maximum (the ties at the maximum sort
to the end in index order and need no sorting network)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_extras import load  # noqa: E402
from qldpc_b200 import Code, graph  # noqa: E402
from qldpc_b200.spaceTime import spaceTimeMatrix  # noqa: E402
H, Lx, d = load("[[144, 12, 12]]")
Hst = spaceTimeMatrix(H, 12)
code = Code(Hst, None, (graph.SEQ, graph.SEQ))
for p, B in ((0.001, 4000), (0.003, 2000), (0.005, 2000)):
    rng = np.random.default_rng(4)
    m, n = H.shape
    err = (rng.random((B, n)) < p).astype(np.int64)
    s = (err @ H.T) % 2
    hist = []
    for _ in range(12):
        s = (s + (rng.random((B, m)) < p)) % 2
        hist.append(s)
    synd = np.concatenate([hist[-1]] + [(hist[i] + hist[i - 1]) % 2 for i in range(1, 12)], axis=1).astype(np.uint8)
    prior = np.full(code.n, np.log((1 - p) / p))
    out = code.bp_decode_batch(synd, prior, variant="min_sum", max_iter=50, alpha=0.8, damping=0.7, clip=25.0, precision=32)
    hard, conv, llr, iters = out
    f = np.nonzero(~np.asarray(conv).astype(bool))[0]
    a = np.abs(np.asarray(llr)[f])
    mx = a.max(axis=1, keepdims=True)
    below = (a < mx).sum(axis=1)
    print(f"p={p}: {len(f)} failed shots; entries below the shot's maximum |LLR|: mean {below.mean():.0f}, median {np.median(below):.0f}, "
          f"90% {np.quantile(below, 0.9):.0f}, max {below.max()}; fraction of shots with <= 1024: {(below <= 1024).mean():.3f}, <= 2048: {(below <= 2048).mean():.3f}; max|llr| median {np.median(mx):.2f}", flush=True)
