"""GPU tests of the rows SURVEY.md section 8f lists as "next": alpha_estimation return paths, the alpha estimator of
rework/Alvarado.py, and the Monte-Carlo drivers that emit the reference's result dictionaries."""
import numpy as np
import pytest

from conftest import load_code_file

pytestmark = pytest.mark.gpu


def _synd(H, errors):
    return ((errors.astype(np.int64) @ (np.asarray(H) != 0).astype(np.int64).T) % 2).astype(np.uint8)


def test_alpha_estimation_return_paths(bp_golden):
    from qldpc_b200.rework import decoding as rw
    d, meta = bp_golden
    for case in meta["cases"]:
        key = case["key"]
        H, _ = load_code_file(case["code"], case["layout"])
        synd = _synd(H, d[key + "_errors"])
        prior = [np.log((1 - case["p"]) / case["p"])] * H.shape[1]
        out = rw.performMinSum_Symmetric(H, synd[0], prior, maxIter=1, alpha=0.8, damping=0.7, clip_llr=25.0, alpha_estimation=True)
        assert out[0] == 0 and out[1] == 0 and out[3] == 0 and out[2].shape == H.shape
        assert np.array_equal(out[2], d[key + "_ms_alphaest"])              # bit-exact
        al, dm, cl = meta["sym_params"]
        out = rw.performBeliefPropagation_Symmetric(H, synd[0], prior, maxIter=50, alpha=al, damping=dm, clip_llr=cl, alpha_estimation=True)
        np.testing.assert_allclose(out[2], d[key + "_sym_alphaest"], rtol=1e-6, atol=1e-9)
    with pytest.raises(ValueError):
        rw.performBeliefPropagation_Symmetric(H, synd[0], prior, maxIter=5, alpha_estimation=True)


def test_estimate_alpha_matches_reference(reference_stats):
    from qldpc_b200.rework.Alvarado import estimate_alpha_from_code
    for rec in reference_stats["alpha_estimates"]:
        H, _ = load_code_file(rec["code"])
        np.random.seed(rec["seed"])
        a = estimate_alpha_from_code(H, trials=rec["trials"], error_rate=rec["p"], maxIter=1, verbose=False)
        assert abs(a - rec["alpha"]) < 1e-9 * max(1.0, abs(rec["alpha"])), (rec, a)


def test_paper_results_dict_and_ler(reference_stats):
    from qldpc_b200 import experiments as X
    p7 = reference_stats["degeneracyCount_p"][7]
    want = reference_stats["BPOSD.npz"]["[[72, 12, 6]]"]["ler"][7]      # sum-product BP50 + OSD-0, single draw
    N = 30000
    res = X.paper_results(codes=["[[72, 12, 6]]"], physicalErrorRates=[p7, 0.01], trials=N, variant="sum_product", maxIter=50,
                          osd_order=0, draws=1, seed=11, precision=64)
    r = res["[[72, 12, 6]]"]
    assert set(r) == {"ler", "BPs_fault", "BPs_miscorrected", "incorrectable", "degeneracies"} and all(len(v) == 2 for v in r.values())
    half = 1.96 * np.sqrt(want * (1 - want) / 10000) + 1.96 * np.sqrt(want * (1 - want) / N)
    assert abs(r["ler"][0] - want) <= half, (r["ler"][0], want, half)
    assert r["BPs_miscorrected"][0] + r["incorrectable"][0] == round(r["ler"][0] * N)
    # BP-only accounting of degeneracyCount.ipynb cell 5
    bp = X.paper_results(codes=["[[72, 12, 6]]"], physicalErrorRates=[p7], trials=N, variant="sum_product", maxIter=50, draws=1,
                         seed=11, precision=64, bp_only=True)["[[72, 12, 6]]"]
    w = reference_stats["BP.npz"]["[[72, 12, 6]]"]
    assert abs(bp["BPs_fault"][0] / N - w["BPs_fault"][7] / 10000) < 0.015
    assert abs(bp["ler"][0] - w["ler"][7]) < 0.03


def test_rework_main_dict(reference_stats):
    from qldpc_b200 import experiments as X
    from qldpc_b200.rework.Alvarado import estimate_alpha_from_code
    exp = [{"code": "[[72, 12, 6]]", "name": "72", "physicalErrorRates": [0.05], "distance": 6}]
    N = 20000
    H, _ = load_code_file("[[72, 12, 6]]")
    alpha = estimate_alpha_from_code(H, trials=5000, error_rate=0.05, maxIter=1, verbose=False, seed=3)
    res = X.rework_main(exp, trials=N, BP_maxIter=50, OSD_order=0, variant="min_sum", alpha=alpha, damping=0.7, clip=25.0, precision=64)
    r = res["72"][0.05]
    assert set(r) == {"logical", "osd", "degeneracies", "average_iterations", "OSD_invocation_AND_logicalError", "weights_found_BP",
                      "weights_found_OSD", "weights_found_BP_error", "weights_found_OSD_error"}
    assert len(r["weights_found_BP_error"]) + len(r["weights_found_OSD_error"]) == round(r["logical"] * N)
    assert len(r["weights_found_OSD_error"]) == round(r["OSD_invocation_AND_logicalError"] * N)
    assert len(r["weights_found_BP"]) + len(r["weights_found_OSD"]) == round(r["degeneracies"] * N)   # all corrections are valid
    # the reference's run of this flow (rework/Alvarado.py -> rework/simulation_results.npz, 10^4 shots): two-sample binomial
    # intervals at z = 3 (three rates compared)
    want = reference_stats["simulation_results.npz"]["72"]["0.05"]
    for key in ("logical", "osd", "degeneracies", "OSD_invocation_AND_logicalError"):
        pooled = (r[key] * N + want[key] * 10000) / (N + 10000)
        assert abs(r[key] - want[key]) <= 3.0 * np.sqrt(pooled * (1 - pooled) * (1 / N + 1 / 10000)), (key, r[key], want[key])
    assert abs(r["average_iterations"] - want["average_iterations"]) < 0.6
    assert min(r["weights_found_BP_error"]) >= 6              # a logical error has weight >= the distance


def test_rework_main_different_orders_dict(tmp_path):
    """rework/main_different_orders.py:45-135 -> simulation_results_complex.npz: results[code][label][p], saved and read back
    the way loadResults.py does (np.load(..., allow_pickle=True)["results"].item())."""
    from qldpc_b200 import experiments as X
    exp = [{"code": "[[72, 12, 6]]", "name": "72", "physicalErrorRates": [0.06, 0.04], "distance": 6}]
    N = 4000
    res = X.rework_main_different_orders(exp, trials=N, variant="min_sum", alpha=0.8, damping=0.7, clip=25.0, precision=64)
    assert list(res) == ["72"] and list(res["72"]) == ["BP50_OSD0", "BP100_OSD0", "BP50_OSD7", "BP100_OSD7"]
    for label, by_p in res["72"].items():
        assert list(by_p) == [0.06, 0.04]
        for r in by_p.values():
            assert set(r) == {"logical", "osd", "degeneracies", "OSD_invocation_AND_logicalError", "weights_found_BP", "weights_found_OSD",
                              "weights_found_BP_error", "weights_found_OSD_error"}
    assert res["72"]["BP100_OSD0"][0.06]["osd"] <= res["72"]["BP50_OSD0"][0.06]["osd"] + 3 * np.sqrt(0.25 / N)   # more iterations, fewer failures
    path = str(tmp_path / "simulation_results_complex.npz")
    X.save_results(path, res)
    f = np.load(path, allow_pickle=True)
    assert list(f.keys()) == ["results"] and f["results"].item() == res


def test_curve_points_use_independent_streams():
    """Every (code, p) point of a driver draws its own range of the Philox stream: the errors at a lower p are not a subset
    of the errors at a higher p (they were, with one shared range per point)."""
    from qldpc_b200 import experiments as X, load_code
    code = load_code("[[72, 12, 6]]")
    a = X.mc_point(code, 0.05, 2000, seed=0, first_shot=0, variant="min_sum", max_iter=20, precision=32)
    b = X.mc_point(code, 0.05, 2000, seed=0, first_shot=2000, variant="min_sum", max_iter=20, precision=32)
    assert a["shots"] == b["shots"] == 2000 and a["error_weight"] != b["error_weight"]
    r1 = X.paper_results(codes=["[[72, 12, 6]]"], physicalErrorRates=[0.05, 0.05], trials=2000, variant="min_sum", maxIter=20, draws=1,
                         precision=32, alpha=0.8, damping=0.7, clip=25.0)["[[72, 12, 6]]"]
    assert r1["incorrectable"][0] != r1["incorrectable"][1] or r1["degeneracies"][0] != r1["degeneracies"][1]


def test_bp_per_iteration_dict():
    from qldpc_b200 import experiments as X
    N = 3000
    res = X.bp_per_iteration(codes=["[[90, 8, 10]]"], errorRate=0.01, iterations=(10, 50), trials=N, variant="sum_product", precision=64)
    r = res["[[90, 8, 10]]"]
    assert set(r) == {"logicalErrors", "degeneracies", "OSD_invocations", "iterations", "llrs_per_iter", "llrs_per_iter_after_OSD"}
    assert r["iterations"] == [10, 50] and r["llrs_per_iter"][0].size == N * 90
    assert r["llrs_per_iter_after_OSD"][1].size == round(r["OSD_invocations"][1] * N) * 90
    assert r["OSD_invocations"][1] <= r["OSD_invocations"][0] + 0.01
    ms = X.bp_per_iteration(codes=["[[108, 8, 10]]"], errorRate=0.01, iterations=(30,), trials=N, variant="min_sum", alpha=0.8,
                            damping=0.7, clip=25.0, precision=32)["[[108, 8, 10]]"]
    assert ms["logicalErrors"][0] < 0.01


def test_device_llr_histograms_match_host_histograms():
    """LLR histograms on the device == np.histogram of the LLRs the batched call returns."""
    from qldpc_b200 import Code, graph
    H, d = load_code_file("[[108, 8, 10]]")
    code = Code(H, d["Lx"], (graph.SEQ, graph.SEQ), int(d["distance"]))
    p, N, seed = 0.04, 30000, 5
    kw = dict(variant="min_sum", max_iter=20, alpha=0.8, damping=0.7, clip=25.0, precision=32)
    h = code.llr_histograms(p, N, lo=-30.0, hi=30.0, nbins=60, seed=seed, **kw)
    err, synd = code.sample(p, N, seed=seed)
    hard, conv, llr, iters = code.bp_decode_batch(synd, np.log((1 - p) / p), **kw)
    edges = h["edges"]
    clipped = np.clip(llr, edges[0], np.nextafter(edges[-1], -np.inf))
    assert np.array_equal(h["true_0"], np.histogram(clipped[err == 0], bins=edges)[0])
    assert np.array_equal(h["true_1"], np.histogram(clipped[err == 1], bins=edges)[0])
    assert np.array_equal(h["bp_failed_shots"], np.histogram(clipped[~conv], bins=edges)[0])
    assert h["n_bp_failed"] == int((~conv).sum()) and int(h["true_0"].sum() + h["true_1"].sum()) == N * 108
