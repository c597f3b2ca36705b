"""Measured accuracy of the sum-product kernels (numbers behind the thresholds of tests/test_gpu_parity.py):
float64 kernels against the reference's golden vectors, float32 psi-domain kernel against the float64 kernel."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_code_file  # noqa: E402
from qldpc_b200 import Code, graph  # noqa: E402

d = np.load(os.path.join(ROOT, "tests", "golden", "bp_golden.npz"))
meta = json.loads(str(d["meta"]))
synd_of = lambda H, e: ((e.astype(np.int64) @ (np.asarray(H) != 0).astype(np.int64).T) % 2).astype(np.uint8)
out = {"f64_vs_reference_golden": [], "f32_psi_vs_f64": []}
for case in meta["cases"]:
    H, _ = load_code_file(case["code"], case["layout"])
    key = case["key"]
    synd = synd_of(H, d[key + "_errors"])
    prior = [np.log((1 - case["p"]) / case["p"])] * H.shape[1]
    code = Code(H, None, graph.reference_schedule(H, "sum_product"))
    hard, conv, llr, iters = code.bp_decode_batch(synd, prior, "sum_product", case["max_iter"], precision=64)
    same = conv & d[key + "_sp_conv"] & (iters == d[key + "_sp_iter"])
    rel = np.abs(llr[same] - d[key + "_sp_llr"][same]) / np.maximum(np.abs(d[key + "_sp_llr"][same]), 1e-300)
    out["f64_vs_reference_golden"].append(dict(case=key, shots=int(len(conv)), flag_agree=float((conv == d[key + "_sp_conv"]).mean()),
                                               both_converged_same_iteration=int(same.sum()), both_converged=int((conv & d[key + "_sp_conv"]).sum()),
                                               hard_identical_on_those=float((hard[same] == d[key + "_sp_hard"][same]).all(1).mean()),
                                               max_rel_llr_err=float(rel.max()) if rel.size else None))
H, _ = load_code_file("[[144, 12, 12]]")
n = H.shape[1]
code = Code(H, None, (graph.SEQ, graph.SEQ))
rng = np.random.default_rng(31)
err = (rng.random((20000, n)) < 0.05).astype(np.uint8)
synd = synd_of(H, err)
prior = [np.log(0.95 / 0.05)] * n
for variant, kw in (("sum_product", {}), ("sum_product_sym", dict(alpha=0.9, damping=0.8, clip=20.0))):
    ref = code.bp_decode_batch(synd, prior, variant, 50, precision=64, **kw)
    got = code.bp_decode_batch(synd, prior, variant, 50, precision=32, **kw)
    same = (got[1] == ref[1]) & (got[3] == ref[3]) & (got[0] == ref[0]).all(1)
    sel = ref[1] & same
    rel = np.abs(got[2][sel] - ref[2][sel]) / np.maximum(np.abs(ref[2][sel]), 1e-3)
    absd = np.abs(got[2][sel] - ref[2][sel])
    out["f32_psi_vs_f64"].append(dict(variant=variant, identical_fraction=float(same.mean()), q99=float(np.quantile(rel, 0.99)), q999=float(np.quantile(rel, 0.999)),
                                      q9999=float(np.quantile(rel, 0.9999)), max_rel=float(rel.max()), max_abs=float(absd.max()),
                                      max_rel_where_abs_llr_ge_1=float(rel[np.abs(ref[2][sel]) >= 1.0].max())))
print(json.dumps(out, indent=1))
