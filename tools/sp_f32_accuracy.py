"""Accuracy of the float32 sum-product kernels against the float64 kernel (which tracks the reference to ~1e-12):
    python tools/sp_f32_accuracy.py [code] [p] [shots]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qldpc_b200 import Code, graph

name = sys.argv[1] if len(sys.argv) > 1 else "[[144, 12, 12]]"
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.05
B = int(sys.argv[3]) if len(sys.argv) > 3 else 20000
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qldpc_b200", "data", "codes", name + ".npz"))
H = d["Hx"]
code = Code(H, d["Lx"], (graph.SEQ, graph.SEQ), int(d["distance"]))
err, synd = code.sample(p, B, seed=3)
prior = np.full(H.shape[1], np.log((1 - p) / p))
for variant, kw in (("sum_product", {}), ("sum_product_sym", dict(alpha=0.9, damping=0.8, clip=20.0))):
    ref = code.bp_decode_batch(synd, prior, variant, 50, precision=64, **kw)
    for label, extra in (("warp psi-domain", {}), ("tiled tanh-domain", dict(lanes_per_shot=8))):
        got = code.bp_decode_batch(synd, prior, variant, 50, precision=32, **extra, **kw)
        kern = code.geometry(code.config(variant, 50, precision=32, **extra, **kw))["kernel"]
        same = (got[1] == ref[1]) & (got[3] == ref[3]) & (got[0] == ref[0]).all(1)
        conv = ref[1] & same
        rel = np.abs(got[2][conv] - ref[2][conv]) / np.maximum(np.abs(ref[2][conv]), 1e-12)
        print(f"{variant:16s} {label:18s} kernel={kern:14s} identical (hard, flag, exit iteration): {same.mean():.5f}   "
              f"LLR rel. error on those converged shots: median {np.median(rel):.2e}  99% {np.quantile(rel, 0.99):.2e}  "
              f"99.99% {np.quantile(rel, 0.9999):.2e}  max {rel.max():.2e}")
