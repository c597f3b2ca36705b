"""One-off throughput probe of a decoder configuration (device-resident inputs, CUDA events).

    python tools/probe.py '{"code": "[[144, 12, 12]]", "p": 0.05, "shots": 2000000, "osd": 7, "cfg": {"variant": "min_sum", ...}}' ...

Each argument is one JSON object; one JSON line is printed per probe ("bp_only": true runs BP without OSD as well)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_extras import Runner, load  # noqa: E402

_runners = {}
for arg in sys.argv[1:]:
    q = json.loads(arg)
    name = q.get("code", "[[144, 12, 12]]")
    if name not in _runners:
        H, Lx, d = load(name)
        _runners[name] = Runner(H, Lx, d)
    r = _runners[name]
    res = r.run(q.get("p", 0.05), q.get("shots", 2_000_000), q["cfg"], q.get("osd", 0), reps=q.get("reps", 2))
    if q.get("bp_only"):
        b = r.run(q.get("p", 0.05), q.get("shots", 2_000_000), q["cfg"], -1, reps=q.get("reps", 2))
        res["bp_only"] = dict(ms=b["ms"], shots_per_s=b["shots_per_s"], shot_iterations_per_s=b["shot_iterations_per_s"])
    print(json.dumps(dict(probe=q, env={k: v for k, v in os.environ.items() if k.startswith("QLDPC_")}, **res)), flush=True)
