"""Per-call latency of the reference-named single-shot wrappers (module-swap use)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qldpc_b200.rework.decoding import performMinSum_Symmetric, performBeliefPropagationFast, performOSD_enhanced
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qldpc_b200", "data", "codes", "[[144, 12, 12]].npz"))
H, n, p = d["Hx"], 144, 0.05
prior = [np.log((1 - p) / p)] * n
rng = np.random.default_rng(0)
errs = (rng.random((300, n)) < p).astype(int)
synd = (errs @ H.T) % 2
for name, fn in (("performMinSum_Symmetric(maxIter=100, 0.8, 0.7, 25)", lambda s: performMinSum_Symmetric(H, s, prior, 100, 0.8, 0.7, 25.0)),
                 ("performBeliefPropagationFast(maxIter=100)", lambda s: performBeliefPropagationFast(H, s, prior, 100))):
    fn(synd[0])
    t0 = time.perf_counter()
    res = [fn(s) for s in synd]
    dt = (time.perf_counter() - t0) / len(synd)
    print(f"{name}: {dt*1e6:.0f} us per call ({sum(not r[1] for r in res)} BP failures of {len(synd)})")
fails = [(s, r) for s, r in zip(synd, res) if not r[1]]
if fails:
    performOSD_enhanced(H, fails[0][0], fails[0][1][2], fails[0][1][0], order=7)      # warm-up: builds the handle
    t0 = time.perf_counter()
    for s, r in fails:
        performOSD_enhanced(H, s, r[2], r[0], order=7)
    print(f"performOSD_enhanced(order=7): {(time.perf_counter() - t0) / len(fails) * 1e6:.0f} us per call")
