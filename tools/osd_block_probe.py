"""BP (CTA-per-shot, float32) + block OSD-0 on the 864 x 2592 space-time matrix at p = 0.001 / 0.003: OSD stage time and rate
(the ncu target of profiles/r2zd_osd_block_ncu_summary.txt: -k regex:osd0_block_fast -c 1 captures the p = 0.001 launch)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_extras import Runner, load  # noqa: E402
from qldpc_b200.spaceTime import spaceTimeMatrix  # noqa: E402
H, Lx, d = load("[[144, 12, 12]]")
r = Runner(spaceTimeMatrix(H, 12))
ms = dict(variant="min_sum", max_iter=50, alpha=0.8, damping=0.7, clip=25.0)
for p, B in ((0.001, 30000), (0.003, 15000)):
    rng = np.random.default_rng(4)
    m, n = H.shape
    err = (rng.random((B, n)) < p).astype(np.int64)
    s = (err @ H.T) % 2
    hist = []
    for _ in range(12):
        s = (s + (rng.random((B, m)) < p)) % 2
        hist.append(s)
    synd = np.concatenate([hist[-1]] + [(hist[i] + hist[i - 1]) % 2 for i in range(1, 12)], axis=1).astype(np.uint8)
    a = r.run(p, B, dict(precision=32, **ms), 0, synd_override=synd, reps=2)
    b = r.run(p, B, dict(precision=32, **ms), -1, synd_override=synd, reps=2)
    osd_ms = a["ms"] - b["ms"]
    nfail = a["bp_failure_rate"] * B
    print(json.dumps(dict(p=p, shots=B, total_ms=a["ms"], bp_ms=b["ms"], osd_ms=osd_ms, bp_failures=nfail, osd_shots_per_s=nfail / osd_ms * 1e3,
                          pipeline_shots_per_s=a["shots_per_s"])), flush=True)
