"""Small run through every kernel, meant to be executed under compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qldpc_b200 import Code, graph
from qldpc_b200.spaceTime import spaceTimeMatrix

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def load(name):
    d = np.load(os.path.join(ROOT, "qldpc_b200", "data", "codes", name + ".npz"))
    return d["Hx"], d["Lx"], int(d["distance"])

rng = np.random.default_rng(0)
for name, B in (("[[72, 12, 6]]", 700), ("[[144, 12, 12]]", 500), ("[[288, 12, 18]]", 200)):
    H, Lx, dist = load(name)
    n = H.shape[1]
    code = Code(H, Lx, graph.reference_schedule(H, "min_sum"), dist)
    err, synd = code.sample(0.06, B, seed=1)
    prior = np.full(n, np.log(0.94 / 0.06))
    for kw in (dict(variant="min_sum", precision=32, alpha=0.8, damping=0.7, clip=25.0),                  # tiled f32 (T auto)
               dict(variant="min_sum", precision=32, alpha=0.8, damping=0.7, clip=25.0, lanes_per_shot=4),
               dict(variant="min_sum", precision=64, alpha=0.8, damping=0.7, clip=25.0),                  # tiled f64
               dict(variant="sum_product", precision=64), dict(variant="sum_product_sym", precision=32, alpha=0.9, damping=0.8),
               dict(variant="min_sum", precision=32, staged=2), dict(variant="min_sum", precision=64, staged=1),
               dict(variant="sum_product", precision=64, staged=2)):
        corr, conv, it = code.bposd_decode_batch(synd, prior, max_iter=20, osd_order=0, **kw)
        chk = code.check_batch(err, corr, synd, conv, it)
        assert chk["valid"].all(), (name, kw)
    hard, conv, llr, it = code.bp_decode_batch(synd, prior, "min_sum", 15, 0.8, 0.7, 25.0, precision=64)
    f = np.nonzero(~conv)[0][:40]
    if f.size:
        code.osd_decode_batch(synd[f], llr[f], hard[f])
        bad = rng.integers(0, 2, (min(4, f.size), H.shape[0])).astype(np.uint8)
        code.osd_decode_batch(bad, llr[f[:len(bad)]], hard[f[:len(bad)]], order=2, max_combinations=40)
    code.mc_sweep(0.05, 3000, seed=2, variant="min_sum", max_iter=30, alpha=0.8, damping=0.7, clip=25.0, osd_order=7)
    code.bp_messages_batch(synd[:8], prior, "min_sum", 1, 1.0, 1.0, 20.0, 0)
    print("ok", name, flush=True)
H, _, _ = load("[[72, 12, 6]]")
Hst = spaceTimeMatrix(H, 12)                    # 432 x 1296: HBM-staged BP + block-per-shot OSD
code = Code(Hst)
synd = (rng.random((60, Hst.shape[0])) < 0.02).astype(np.uint8)
prior = np.full(Hst.shape[1], np.log(0.99 / 0.01))
corr, conv, it = code.bposd_decode_batch(synd, prior, "min_sum", 12, 0.8, 0.7, 25.0, precision=32, osd_order=0)
assert (((corr.astype(np.int64) @ (Hst != 0).astype(np.int64).T) % 2) == synd).all()
print("ok space-time", flush=True)
