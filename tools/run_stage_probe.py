"""Small fixed workload on the 864 x 2592 space-time matrix for ncu captures: the staged CTA kernel in float32 (BP only), then
float64 min-sum BP + block OSD-0, then float64 sum-product BP only (one launch each)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_extras import Runner, load  # noqa: E402
from qldpc_b200.spaceTime import spaceTimeMatrix  # noqa: E402

H, Lx, d = load("[[144, 12, 12]]")
r = Runner(spaceTimeMatrix(H, 12))
p, B = 0.003, 40000
rng = np.random.default_rng(4)
m, n = H.shape
err = (rng.random((B, n)) < p).astype(np.int64)
s = (err @ H.T) % 2
hist = []
for _ in range(12):
    s = (s + (rng.random((B, m)) < p)) % 2
    hist.append(s)
synd = np.concatenate([hist[-1]] + [(hist[i] + hist[i - 1]) % 2 for i in range(1, 12)], axis=1).astype(np.uint8)
ms = dict(variant="min_sum", max_iter=50, alpha=0.8, damping=0.7, clip=25.0)
out = {}
out["f32_staged_bp_only"] = r.run(p, B, dict(staged=5, precision=32, **ms), -1, synd_override=synd, reps=1)
out["f64_bp_osd"] = r.run(p, B // 2, dict(precision=64, **ms), 0, synd_override=synd[:B // 2], reps=1)
out["f64_sum_product_bp_only"] = r.run(p, B // 4, dict(variant="sum_product", max_iter=50, precision=64), -1, synd_override=synd[:B // 4], reps=1)
print(json.dumps(out))
