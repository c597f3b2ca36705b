"""CPU tests: the oracle (oracle/) against the golden vectors produced by the unmodified
reference (tools/make_golden.py).  This is what "pins" the oracle."""
import numpy as np
import pytest

from conftest import load_code_file
from oracle import oracle as O


def _synd(H, errors):
    return (errors.astype(np.int64) @ np.asarray(H).T) % 2


def _prior(p, n):
    return [np.log((1 - p) / p)] * n


@pytest.mark.parametrize("ci", range(6))
def test_min_sum_bit_exact(bp_golden, ci):
    d, meta = bp_golden
    case = meta["cases"][ci]
    H, _ = load_code_file(case["code"], case["layout"])
    key = case["key"]
    errors = d[key + "_errors"]
    synd = _synd(H, errors)
    g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))
    prior = _prior(case["p"], H.shape[1])
    for pi, (al, dm, cl) in enumerate(meta["minsum_params"]):
        for i in range(case["shots"]):
            hard, ok, llr, it = O.bp_decode(g, synd[i], prior, O.MIN_SUM, case["max_iter"], al, dm, cl)
            assert np.array_equal(hard, d[f"{key}_ms{pi}_hard"][i])
            assert ok == bool(d[f"{key}_ms{pi}_conv"][i])
            assert it == int(d[f"{key}_ms{pi}_iter"][i])
            assert np.array_equal(llr, d[f"{key}_ms{pi}_llr"][i])  # float64 bit-exact


@pytest.mark.parametrize("ci", range(6))
def test_sum_product_numpy_bit_exact(bp_golden, ci):
    d, meta = bp_golden
    case = meta["cases"][ci]
    H, _ = load_code_file(case["code"], case["layout"])
    key = case["key"]
    synd = _synd(H, d[key + "_errors"])
    g = O.Graph(H, *O.auto_schedule(H, O.SUM_PRODUCT))
    gs = O.Graph(H, *O.auto_schedule(H, O.SUM_PRODUCT_SYM))
    prior = _prior(case["p"], H.shape[1])
    al, dm, cl = meta["sym_params"]
    for i in range(case["shots"]):
        hard, ok, llr, it = O.sum_product_numpy(g, synd[i], prior, case["max_iter"])
        assert np.array_equal(hard, d[key + "_sp_hard"][i]) and ok == bool(d[key + "_sp_conv"][i])
        assert it == int(d[key + "_sp_iter"][i])
        assert np.array_equal(llr, d[key + "_sp_llr"][i])
        hard, ok, llr, it = O.sum_product_numpy(gs, synd[i], prior, case["max_iter"], al, dm, cl)
        assert np.array_equal(hard, d[key + "_sym_hard"][i]) and ok == bool(d[key + "_sym_conv"][i])
        assert it == int(d[key + "_sym_iter"][i])
        assert np.array_equal(llr, d[key + "_sym_llr"][i])


@pytest.mark.parametrize("ci", range(6))
def test_sum_product_c_close(bp_golden, ci):
    """glibc tanh/atanh differ from NumPy's in the last ulp: compare shots that converge early
    (no chaotic amplification) tightly, and require the rest to agree on the convergence flag mostly."""
    d, meta = bp_golden
    case = meta["cases"][ci]
    H, _ = load_code_file(case["code"], case["layout"])
    key = case["key"]
    synd = _synd(H, d[key + "_errors"])
    g = O.Graph(H, *O.auto_schedule(H, O.SUM_PRODUCT))
    prior = _prior(case["p"], H.shape[1])
    n_checked = 0
    for i in range(case["shots"]):
        hard, ok, llr, it = O.bp_decode(g, synd[i], prior, O.SUM_PRODUCT, case["max_iter"])
        if d[key + "_sp_conv"][i] and d[key + "_sp_iter"][i] <= 15:
            assert ok and it == int(d[key + "_sp_iter"][i])
            assert np.array_equal(hard, d[key + "_sp_hard"][i])
            np.testing.assert_allclose(llr, d[key + "_sp_llr"][i], rtol=1e-9)
            n_checked += 1
    assert n_checked >= case["shots"] // 2


def test_loop_version_and_alpha_estimation(bp_golden):
    d, meta = bp_golden
    for case in meta["cases"]:
        key = case["key"]
        H, _ = load_code_file(case["code"], case["layout"])
        synd = _synd(H, d[key + "_errors"])
        prior = _prior(case["p"], H.shape[1])
        if key + "_loop_hard" in d.files:
            g = O.Graph(H, O.SEQ, O.SEQ)  # loop version sums a gathered vector sequentially
            for i in range(len(d[key + "_loop_hard"])):
                hard, ok, llr, it = O.sum_product_numpy(g, synd[i], prior, case["max_iter"])
                assert np.array_equal(hard, d[key + "_loop_hard"][i]) and ok == bool(d[key + "_loop_conv"][i])
                assert np.array_equal(llr, d[key + "_loop_llr"][i])
        g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))
        r = O.bp_alpha_messages(g, synd[0], prior, O.MIN_SUM, 1, 0.8, 0.7, 25.0)
        assert np.array_equal(r, d[key + "_ms_alphaest"])
        al, dm, cl = meta["sym_params"]
        r = O.bp_alpha_messages(g, synd[0], prior, O.SUM_PRODUCT_SYM, 50, al, dm, cl)
        np.testing.assert_allclose(r, d[key + "_sym_alphaest"], rtol=1e-6, atol=1e-9)


def test_non_uniform_prior(bp_golden):
    d, _ = bp_golden
    H, _ = load_code_file("[[72, 12, 6]]")
    prior = d["nu_prior"]
    synd = _synd(H, d["nu_errors"])
    g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))
    gsp = O.Graph(H, *O.auto_schedule(H, O.SUM_PRODUCT))
    for i in range(len(synd)):
        hard, ok, llr, it = O.bp_decode(g, synd[i], prior, O.MIN_SUM, 30, 0.75, 0.7, 25.0)
        assert np.array_equal(hard, d["nu_ms_hard"][i]) and ok == bool(d["nu_ms_conv"][i]) and it == d["nu_ms_iter"][i]
        assert np.array_equal(llr, d["nu_ms_llr"][i])
        hard, ok, llr, it = O.sum_product_numpy(gsp, synd[i], prior, 30)
        assert np.array_equal(hard, d["nu_sp_hard"][i]) and ok == bool(d["nu_sp_conv"][i])
        assert np.array_equal(llr, d["nu_sp_llr"][i])


def test_osd_after_bp_bit_exact(bp_golden):
    """OSD-0 on the BP-failed golden shots with the BP LLRs as they are (ties included): the
    reference was fed the stable ranks (SURVEY.md H1 contract)."""
    d, meta = bp_golden
    for case in meta["cases"]:
        key = case["key"]
        H, _ = load_code_file(case["code"], case["layout"])
        synd = _synd(H, d[key + "_errors"])
        g = O.Graph(H)
        for pi in range(2):
            for i in np.nonzero(~d[f"{key}_ms{pi}_conv"])[0]:
                sol = O.osd0(g, synd[i], d[f"{key}_ms{pi}_llr"][i], d[f"{key}_ms{pi}_hard"][i])
                assert np.array_equal(sol, d[f"{key}_ms{pi}_osd0"][i])
                sol7, swept = O.osd_enhanced(g, synd[i], d[f"{key}_ms{pi}_llr"][i], d[f"{key}_ms{pi}_hard"][i], order=7)
                assert np.array_equal(sol7, sol) and not swept  # H5: OSD-w == OSD-0 on consistent syndromes


def test_osd_golden(osd_golden):
    d, meta = osd_golden
    for case in meta["cases"]:
        key = case["key"]
        H, _ = load_code_file(case["code"])
        g = O.Graph(H)
        llr, hard = d[key + "_llr"], d[key + "_hard"]
        for i in range(case["shots"]):
            assert np.array_equal(O.osd0(g, d[key + "_synd_c"][i], llr[i], hard[i]), d[key + "_osd0_c"][i])
            assert np.array_equal(O.osd0(g, d[key + "_synd_i"][i], llr[i], hard[i]), d[key + "_osd0_i"][i])
            sol, swept = O.osd_enhanced(g, d[key + "_synd_c"][i], llr[i], hard[i], order=7)
            assert np.array_equal(sol, d[key + "_enh7_c"][i]) and not swept
        for order, mc in case["sweeps"]:
            ref = d[f"{key}_enh_o{order}_mc{mc}"]
            for i in range(len(ref)):
                sol, _ = O.osd_enhanced(g, d[key + "_synd_i"][i], d[key + "_llr_tf"][i], hard[i], order=order,
                                        max_combinations=(mc or None))
                assert np.array_equal(sol, ref[i]), (case["code"], order, mc, i)


def test_spacetime(spacetime_golden):
    d = spacetime_golden
    H, _ = load_code_file("[[72, 12, 6]]")
    Hst = O.space_time_matrix(H, 3)
    ref = np.zeros(tuple(d["Hst_shape"]))
    ref[d["Hst_rows"], d["Hst_cols"]] = 1.0
    assert Hst.dtype == np.float64 and Hst.flags["C_CONTIGUOUS"] and np.array_equal(Hst, ref)
    np.random.seed(5)
    e, s = O.spacetime_syndrome(H, 0.03, 3)
    assert np.array_equal(e, d["seed5_error"]) and np.array_equal(s, d["seed5_syndrome"])
    g = O.Graph(Hst, *O.auto_schedule(Hst, O.MIN_SUM))
    p = 0.02
    prior = _prior(p, Hst.shape[1])
    for i in range(len(d["synd"])):
        hard, ok, llr, it = O.bp_decode(g, d["synd"][i], prior, O.MIN_SUM, 50, 0.8, 0.7, 25.0)
        assert np.array_equal(hard, d["ms_hard"][i]) and ok == bool(d["ms_conv"][i]) and it == d["ms_iter"][i]
        assert np.array_equal(llr, d["ms_llr"][i])
        hard2, ok2, llr2, _ = O.sum_product_numpy(g, d["synd"][i], prior, 50)
        assert np.array_equal(hard2, d["sp_hard"][i]) and ok2 == bool(d["sp_conv"][i]) and np.array_equal(llr2, d["sp_llr"][i])
        if not ok:
            assert np.array_equal(O.osd0(g, d["synd"][i], llr, hard), d["ms_osd0"][i])


def test_kat_bp_npz(reference_stats):
    """notebooks/data/BP.npz, [[72,12,6]] row: RNG -> syndrome -> sum-product BP(50) -> flag -> logical
    check, seed 0.  Replays p-points 3..6 exactly (10 000 shots each) with the C
    oracle; the RNG stream of the skipped points is consumed without decoding."""
    st = reference_stats
    ps = np.logspace(-3.2, -1.3, 8)
    H, d = load_code_file("[[72, 12, 6]]")
    Lx, dist = d["Lx"], int(d["distance"])
    g = O.Graph(H, *O.auto_schedule(H, O.SUM_PRODUCT))
    n = H.shape[1]
    want = st["BP.npz"]["[[72, 12, 6]]"]
    np.random.seed(0)
    for pi, p in enumerate(ps):
        errors = np.array([(np.random.random(n) < p) for _ in range(10000)], np.uint8)
        if pi not in (3, 4, 5, 6):
            continue
        synd = _synd(H, errors)
        r = O.decode_batch(g, synd, _prior(p, n), O.SUM_PRODUCT, 50, osd_order=-1)
        logical, valid, _ = O.check_batch(g, Lx, errors, r["corr"], synd)
        bp_fault = int((~r["converged"]).sum())
        wt = errors.sum(1)
        conv = r["converged"]
        # degeneracyCount.ipynb cell 5: a BP failure counts as a logical error AND the logical check of the
        # (non-converged) detection is counted on top (the double-counting noted in SURVEY.md section 6)
        ler = (bp_fault + int(logical.sum())) / 10000
        incorrectable = int((logical & (wt >= dist // 2)).sum())
        assert int((logical & (wt < dist // 2)).sum()) == want["BPs_miscorrected"][pi]
        degener = int((conv & ~logical & (r["corr"] != errors).any(1)).sum())
        assert bp_fault == want["BPs_fault"][pi]
        assert degener == want["degeneracies"][pi]
        assert abs(ler - want["ler"][pi]) < 1e-12, (pi, ler, want["ler"][pi])
        assert incorrectable == want["incorrectable"][pi]
