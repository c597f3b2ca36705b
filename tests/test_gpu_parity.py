"""GPU parity tests (run on the B200 box with -m gpu): the CUDA path, called through the C ABI
(qldpc_b200.Code -> libqldpc_b200.so), against the golden vectors of the unmodified reference and
against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): OSD / syndrome / logical checks bit-exact; float64 min-sum
bit-exact including LLRs and exit iteration; float32 min-sum equal hard decisions except on
trajectories that have already diverged (reported); sum-product LLRs within 1e-4 relative on shots
that converge at the same iteration.
"""
import numpy as np
import pytest

from conftest import load_code_file
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _synd(H, errors):
    return ((errors.astype(np.int64) @ (np.asarray(H) != 0).astype(np.int64).T) % 2).astype(np.uint8)


def _prior(p, n):
    return [np.log((1 - p) / p)] * n


def _code(H, variant, L=None, distance=None):
    from qldpc_b200 import Code, graph
    return Code(H, L, graph.reference_schedule(H, variant), distance)


# ----------------------------------------------------------------------------------------------
# BP
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ci", range(6))
def test_min_sum_f64_bit_exact_vs_reference_golden(bp_golden, ci):
    d, meta = bp_golden
    case = meta["cases"][ci]
    H, _ = load_code_file(case["code"], case["layout"])
    key = case["key"]
    synd = _synd(H, d[key + "_errors"])
    code = _code(H, "min_sum")
    prior = _prior(case["p"], H.shape[1])
    for pi, (al, dm, cl) in enumerate(meta["minsum_params"]):
        hard, conv, llr, iters = code.bp_decode_batch(synd, prior, "min_sum", case["max_iter"], al, dm, cl, precision=64)
        assert np.array_equal(hard, d[f"{key}_ms{pi}_hard"])
        assert np.array_equal(conv, d[f"{key}_ms{pi}_conv"])
        assert np.array_equal(iters, d[f"{key}_ms{pi}_iter"])
        assert np.array_equal(llr, d[f"{key}_ms{pi}_llr"])          # float64, bit for bit


def test_reference_named_wrappers(bp_golden):
    """The drop-in function names, with the reference's dtypes and tuple shapes."""
    from qldpc_b200.rework import decoding as rw
    from qldpc_b200.decoding import beliefPropagation as bpm, OSD as osdm, OSD_enhanced as enh, beliefPropagationGPU as gpu
    d, meta = bp_golden
    case = meta["cases"][0]
    H, _ = load_code_file(case["code"], case["layout"])
    key = case["key"]
    synd = _synd(H, d[key + "_errors"])
    prior = _prior(case["p"], H.shape[1])
    al, dm, cl = meta["minsum_params"][1]
    for i in range(6):
        hard, ok, llr, it = rw.performMinSum_Symmetric(H, synd[i].astype(np.int64), prior, maxIter=case["max_iter"], alpha=al,
                                                       damping=dm, clip_llr=cl)
        assert hard.dtype == np.int8 and isinstance(ok, bool) and llr.dtype == np.float64 and isinstance(it, int)
        assert np.array_equal(hard, d[f"{key}_ms1_hard"][i]) and ok == bool(d[f"{key}_ms1_conv"][i])
        assert it == d[f"{key}_ms1_iter"][i] and np.array_equal(llr, d[f"{key}_ms1_llr"][i])
        out = bpm.performBeliefPropagationFast(H, synd[i], prior, verbose=False, maxIter=case["max_iter"])
        assert len(out) == 3 and out[0].dtype == np.int8 and out[2].dtype == np.float64
        out4 = rw.performBeliefPropagationFast(H, synd[i], prior, maxIter=case["max_iter"])
        assert len(out4) == 4 and np.array_equal(out4[0], out[0])
        if d[key + "_sp_conv"][i] and d[key + "_sp_iter"][i] <= 15:
            assert np.array_equal(out[0], d[key + "_sp_hard"][i]) and out[1]
            np.testing.assert_allclose(out[2], d[key + "_sp_llr"][i], rtol=1e-6)
        if not d[f"{key}_ms0_conv"][i]:
            sol = osdm.performOSD(H, synd[i], d[f"{key}_ms0_llr"][i], d[f"{key}_ms0_hard"][i])
            assert sol.dtype == np.int64 and np.array_equal(sol, d[f"{key}_ms0_osd0"][i])
            sol7 = enh.performOSD_enhanced(H, synd[i], d[f"{key}_ms0_llr"][i], d[f"{key}_ms0_hard"][i], order=7)
            assert np.array_equal(sol7, sol)
            sol7b = rw.performOSD_enhanced(H.astype(np.float64), synd[i], d[f"{key}_ms0_llr"][i], d[f"{key}_ms0_hard"][i], order=7)
            assert np.array_equal(sol7b, sol)       # float H accepted (the reference raises TypeError)
    hb, cb, lb = gpu.performBeliefPropagationBatch(H, synd, prior, maxIter=case["max_iter"])
    assert hb.dtype == np.int8 and cb.dtype == bool and lb.dtype == np.float64 and hb.shape == (len(synd), H.shape[1])
    e, s = gpu.generate_errors_and_syndromes_batch(H, 0.05, 100, np.random.default_rng(0))
    e2 = (np.random.default_rng(0).random((100, H.shape[1])) < 0.05).astype(np.int8)
    assert np.array_equal(e, e2) and np.array_equal(s, _synd(H, e2)) and s.dtype == np.int8
    e3, s3 = gpu.generate_errors_and_syndromes_batch(H, 0.05, 100)
    assert np.array_equal(s3.astype(np.uint8), _synd(H, e3))


@pytest.mark.parametrize("ci", range(6))
def test_sum_product_f64_vs_reference_golden(bp_golden, ci):
    """CUDA's tanh/atanh are not NumPy's (last-ulp differences), so sum-product parity is a tolerance: the north star asks for
    1e-4 relative.  Measured against the unmodified reference's golden vectors (tools/sp_accuracy_report.py ->
    profiles/r2_sp_accuracy.json): identical flags, exit iterations and hard decisions on every golden shot, LLRs within 3.2e-9
    relative.  Held here to 1e-7 on EVERY shot that converges (not only the early ones) and to exact agreement of the rest."""
    d, meta = bp_golden
    case = meta["cases"][ci]
    H, _ = load_code_file(case["code"], case["layout"])
    key = case["key"]
    synd = _synd(H, d[key + "_errors"])
    prior = _prior(case["p"], H.shape[1])
    code = _code(H, "sum_product")
    hard, conv, llr, iters = code.bp_decode_batch(synd, prior, "sum_product", case["max_iter"], precision=64)
    assert np.array_equal(conv, d[key + "_sp_conv"]) and np.array_equal(iters, d[key + "_sp_iter"])
    ok = d[key + "_sp_conv"]
    assert ok.sum() >= case["shots"] // 2
    assert np.array_equal(hard[ok], d[key + "_sp_hard"][ok])
    np.testing.assert_allclose(llr[ok], d[key + "_sp_llr"][ok], rtol=1e-7)
    # damped / scaled / clipped variant (rework/decoding.py:131)
    al, dm, cl = meta["sym_params"]
    codes = _code(H, "sum_product_sym")
    hard, conv, llr, iters = codes.bp_decode_batch(synd, prior, "sum_product_sym", case["max_iter"], al, dm, cl, precision=64)
    ok = d[key + "_sym_conv"]
    assert np.array_equal(conv, d[key + "_sym_conv"])
    assert np.array_equal(iters[ok], d[key + "_sym_iter"][ok]) and np.array_equal(hard[ok], d[key + "_sym_hard"][ok])
    np.testing.assert_allclose(llr[ok], d[key + "_sym_llr"][ok], rtol=1e-7)


def test_sum_product_f32_accuracy_is_measured_not_assumed(bp_golden):
    """float32 sum-product (fast variant): SURVEY.md H4 predicts it cannot meet 1e-4 relative everywhere (the check-node
    clip 0.9999999 is not representable); it must still agree on hard decisions for early-converging shots and its LLR
    error is printed, bounded loosely.  The float64 kernels are the ones held to the 1e-4 bar."""
    d, meta = bp_golden
    worst = 0.0
    for case in meta["cases"]:
        key = case["key"]
        H, _ = load_code_file(case["code"], case["layout"])
        synd = _synd(H, d[key + "_errors"])
        prior = _prior(case["p"], H.shape[1])
        code = _code(H, "sum_product")
        hard, conv, llr, iters = code.bp_decode_batch(synd, prior, "sum_product", case["max_iter"], precision=32)
        early = d[key + "_sp_conv"] & (d[key + "_sp_iter"] <= 10)
        agree = (hard[early] == d[key + "_sp_hard"][early]).all(1).mean()
        assert agree >= 0.9
        same = early & conv & (iters == d[key + "_sp_iter"])
        rel = np.abs(llr[same] - d[key + "_sp_llr"][same]) / np.maximum(1e-3, np.abs(d[key + "_sp_llr"][same]))
        worst = max(worst, float(rel.max()))
    print(f"\n[f32 sum-product] worst relative LLR error on shots converging at the reference's iteration: {worst:.2e}")
    assert worst < 0.05


def test_sum_product_f32_psi_domain_kernel_vs_f64():
    """The float32 sum-product of the warp-per-shot kernel works in the psi domain (psi(a) = -log tanh(a/2)), which keeps
    float32's relative accuracy near the check-node saturation where the tanh domain cannot: against the float64 kernel
    it must (almost always) take the same decisions and exit at the same iteration, with a 99th-percentile relative LLR
    error below 1e-4, and beat the tanh-domain float32 kernel on both counts."""
    H, _ = load_code_file("[[144, 12, 12]]")
    n = H.shape[1]
    from qldpc_b200 import Code, graph
    code = Code(H, None, (graph.SEQ, graph.SEQ))
    rng = np.random.default_rng(31)
    err = (rng.random((6000, n)) < 0.05).astype(np.uint8)
    synd = _synd(H, err)
    prior = _prior(0.05, n)
    for variant, kw in (("sum_product", {}), ("sum_product_sym", dict(alpha=0.9, damping=0.8, clip=20.0))):
        ref = code.bp_decode_batch(synd, prior, variant, 50, precision=64, **kw)
        stats = {}
        for label, extra in (("psi", {}), ("tanh", dict(lanes_per_shot=8))):
            cfg = code.config(variant, 50, precision=32, **extra, **kw)
            assert code.geometry(cfg)["kernel"] == ("warp_per_shot" if label == "psi" else "tiled")
            got = code.bp_decode_batch(synd, prior, variant, 50, precision=32, **extra, **kw)
            same = (got[1] == ref[1]) & (got[3] == ref[3]) & (got[0] == ref[0]).all(1)
            sel = ref[1] & same
            rel = np.abs(got[2][sel] - ref[2][sel]) / np.maximum(np.abs(ref[2][sel]), 1e-3)
            stats[label] = (same.mean(), np.quantile(rel, 0.99), np.quantile(rel, 0.999), rel.max())
        print(f"\n[f32 {variant}] identical fraction / q99 / q99.9 / max rel. LLR error: psi {stats['psi']}, tanh {stats['tanh']}")
        # The bound is a DISTRIBUTION, not a maximum (profiles/r2_sp_accuracy.json, 20 000 shots: q99 4.6e-5, q99.9
        # 4.8e-3, max 5.8 relative on a posterior near zero after ~40 iterations): float32 trajectories of slowly converging shots
        # drift like the min-sum ones do.  The float64 kernels carry the 1e-4 bar (test_sum_product_f64_vs_reference_golden).
        assert stats["psi"][0] >= 0.99 and stats["psi"][1] < 1e-4 and stats["psi"][2] < 2e-2, stats
        assert stats["psi"][0] >= stats["tanh"][0] - 0.002 and stats["psi"][1] <= stats["tanh"][1], stats


@pytest.mark.parametrize("stem,p", [("[[72, 12, 6]]", 0.06), ("[[90, 8, 10]]", 0.05), ("[[108, 8, 10]]", 0.05),
                                    ("[[144, 12, 12]]", 0.05), ("[[288, 12, 18]]", 0.07)])
def test_sum_product_f64_warp_kernel_vs_tiled_kernel(stem, p):
    """The warp-per-shot float64 sum-product kernels (bp_warp_kernel_f64, VAR 1 / 2: own tanh / atanh, leave-one-out products)
    against the tiled float64 kernel (math-library tanh / atanh, the reference's division by the own factor): same flags,
    exit iterations and hard decisions, LLRs within 1e-9 relative on the shots that converge early and 1e-4 on all -- uniform priors, non-uniform
    priors (the two-table instantiation on the Fortran-ordered Hx) and a prior of exactly 0, which puts tanh(Q/2) = 0 factors
    into iteration 0 and sends their checks through the reference's `tanh_Q_safe` division path."""
    H, _ = load_code_file(stem)
    n = H.shape[1]
    rng = np.random.default_rng(7)
    B = 1501
    synd = _synd(H, (rng.random((B, n)) < p).astype(np.uint8))
    pri_nu = np.log((1 - p) / p) * rng.uniform(0.6, 1.4, n)
    pri_zero = pri_nu.copy()
    pri_zero[rng.choice(n, 3, replace=False)] = 0.0
    for variant, kw in (("sum_product", {}), ("sum_product_sym", dict(alpha=0.9, damping=0.8, clip=20.0))):
        code = _code(H, variant)
        for prior in (_prior(p, n), pri_nu, pri_zero):
            assert code.geometry(code.config(variant, 40, precision=64, **kw))["kernel"] == "warp_per_shot"
            assert code.geometry(code.config(variant, 40, precision=64, lanes_per_shot=8, **kw))["kernel"] == "tiled"
            got = code.bp_decode_batch(synd, prior, variant, 40, precision=64, **kw)
            ref = code.bp_decode_batch(synd, prior, variant, 40, precision=64, lanes_per_shot=8, **kw)
            same = (got[1] == ref[1]) & (got[3] == ref[3]) & (got[0] == ref[0]).all(1)
            assert same.mean() >= 0.999, (variant, same.mean())
            sel = same & ref[1]
            assert sel.sum() > B // 3
            rel = (np.abs(got[2] - ref[2]) / np.maximum(np.abs(ref[2]), 1e-3)).max(1)
            early = sel & (ref[3] < 12)
            worst = int(np.argmax(np.where(sel, rel, 0)))
            print(f"\n[f64 {variant} warp vs tiled, {stem}] identical {same.mean():.4f}; max rel. LLR error: exit iteration < 12 "
                  f"{rel[early].max():.2e}, all {rel[sel].max():.2e} (a shot that exits at iteration {ref[3][worst]}), q99.9 {np.quantile(rel[sel], 0.999):.2e}")
            # Undamped sum-product amplifies a last-place difference by ~2x per iteration on the shots that wander before they
            # converge (measured maxima by exit iteration: < 5: 2e-11, < 10: 3e-10, < 20: 1e-7, < 30: 1e-6, < 40: 5.5e-5 -- a [[108,8,10]]
            # shot that exits at iteration 39; flags, iterations and hard decisions identical on every shot): the early shots carry
            # the tight bar, every shot the north-star one.
            print("   by exit iteration (<5, <10, <20, <30, <40):", ["%.1e" % rel[sel & (ref[3] >= a) & (ref[3] < b)].max(initial=0) for a, b in ((0, 5), (5, 10), (10, 20), (20, 30), (30, 40))])
            assert rel[early].max() < 1e-9 and rel[sel].max() < 1e-4 and np.quantile(rel[sel], 0.99) < 1e-6, (variant, rel[sel].max())


def test_non_uniform_prior_and_loop_version(bp_golden):
    d, meta = bp_golden
    H, _ = load_code_file("[[72, 12, 6]]")
    synd = _synd(H, d["nu_errors"])
    code = _code(H, "min_sum")
    hard, conv, llr, iters = code.bp_decode_batch(synd, d["nu_prior"], "min_sum", 30, 0.75, 0.7, 25.0, precision=64)
    assert np.array_equal(hard, d["nu_ms_hard"]) and np.array_equal(conv, d["nu_ms_conv"])
    assert np.array_equal(iters, d["nu_ms_iter"]) and np.array_equal(llr, d["nu_ms_llr"])
    # loop version: sequential sums
    from qldpc_b200.decoding.beliefPropagation import performBeliefPropagation
    case = meta["cases"][0]
    s2 = _synd(H, d["c0_errors"])
    for i in range(len(d["c0_loop_hard"])):
        if d["c0_loop_conv"][i]:
            h, ok, l = performBeliefPropagation(H, s2[i], _prior(case["p"], 72), verbose=False, maxIter=case["max_iter"])
            assert ok and np.array_equal(h, d["c0_loop_hard"][i])
            np.testing.assert_allclose(l, d["c0_loop_llr"][i], rtol=1e-6)


@pytest.mark.parametrize("stem,p,params", [("[[72, 12, 6]]", 0.05, (1.0, 1.0, 20.0)), ("[[72, 12, 6]]", 0.05, (0.8, 0.7, 25.0)),
                                           ("[[144, 12, 12]]", 0.05, (0.8, 0.7, 25.0)), ("[[288, 12, 18]]", 0.06, (0.8, 0.7, 25.0))])
def test_min_sum_f32_vs_f64_oracle(stem, p, params):
    """float32 production kernel against the float64 oracle on 3000 seeded shots.  BP on non-converging
    shots is chaotic: shots that converge in both precisions must give the same hard
    decision apart from rare rounding-induced path changes; the disagreement rate is bounded and printed."""
    H, _ = load_code_file(stem)
    n = H.shape[1]
    rng = np.random.default_rng(11)
    err = (rng.random((3000, n)) < p).astype(np.uint8)
    synd = _synd(H, err)
    al, dm, cl = params
    g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))
    ref = O.decode_batch(g, synd, _prior(p, n), O.MIN_SUM, 50, al, dm, cl, osd_order=-1, want_llr=True)
    code = _code(H, "min_sum")
    hard, conv, llr, iters = code.bp_decode_batch(synd, _prior(p, n), "min_sum", 50, al, dm, cl, precision=32)
    both = conv & ref["converged"]
    same_flag = (conv == ref["converged"]).mean()
    same_hard = (hard[both].astype(np.uint8) == ref["corr"][both]).all(1).mean()
    same_iter = (iters[both] == ref["iters"][both]).mean()
    rel = np.abs(llr[both] - ref["llr"][both]) / np.maximum(1e-6, np.abs(ref["llr"][both]))
    print(f"\n[f32 vs f64 oracle] {stem} {params}: same conv flag {same_flag:.4f}, same hard (both converged) {same_hard:.4f}, "
          f"same exit iter {same_iter:.4f}, median LLR rel err {np.median(rel):.2e}")
    assert same_flag >= 0.97 and same_hard >= 0.97 and same_iter >= 0.95
    # every converged float32 decision really satisfies its syndrome
    assert np.array_equal(_synd(H, hard[conv].astype(np.uint8)), synd[conv])


def test_f32_divergence_is_iteration_resolved():
    """North-star criterion for float32 min-sum: hard decisions equal the float64 reference's, mismatches confined to
    near-ties.  BP on non-converging shots is chaotic, so the comparison is made iteration by iteration: for
    every shot whose float32 and float64 hard decisions ever differ, find the FIRST iteration where they do and look at the
    float64 |LLR| of the differing bits there.  The fraction explained by |LLR| < 1e-5 is printed, not assumed; what is
    asserted is that divergence is rare, starts late, and never touches shots that converge in float64 within the window."""
    H, _ = load_code_file("[[144, 12, 12]]")
    n, p = 144, 0.05
    rng = np.random.default_rng(31)
    err = (rng.random((8000, n)) < p).astype(np.uint8)
    synd = _synd(H, err)
    prior = _prior(p, n)
    code = _code(H, "min_sum")
    K = 60
    first = np.full(len(synd), -1)
    margin = np.full(len(synd), np.nan)
    conv64_at = np.full(len(synd), 10**9)
    for k in range(1, K + 1):
        h32, c32, l32, i32 = code.bp_decode_batch(synd, prior, "min_sum", k, 0.8, 0.7, 25.0, precision=32)
        h64, c64, l64, i64 = code.bp_decode_batch(synd, prior, "min_sum", k, 0.8, 0.7, 25.0, precision=64)
        conv64_at = np.where(c64 & (conv64_at > K), i64, conv64_at)
        live = ~(c32 & c64 & (i32 == i64))               # frozen (converged at the same iteration) shots cannot change any more
        diff = (h32 != h64)
        newly = diff.any(1) & (first < 0) & live
        first[newly] = k - 1
        for s in np.nonzero(newly)[0]:
            margin[s] = np.abs(l64[s][diff[s]]).min()
    bad = first >= 0
    nbad = int(bad.sum())
    tie = int((margin[bad] < 1e-5).sum())
    small = int((margin[bad] < 1e-3 * prior[0]).sum())
    print(f"\n[f32 vs f64, iteration-resolved] {nbad}/{len(synd)} shots ever differ in a hard decision within {K} iterations; "
          f"first difference at |LLR64| < 1e-5: {tie}, < 1e-3*prior: {small}; max margin {np.nanmax(margin) if nbad else 0:.3g}")
    # measured on B200 (8000 shots, 60 iterations): a handful of shots, all non-converging for >= 30 iterations, where
    # rounding differences accumulate before any hard decision flips (margins up to ~0.4, i.e. not ties): reported as the
    # "unexplained" fraction SURVEY.md H3 predicts.  Hard guarantees: rare, and never on a shot float64 has already decoded.
    assert nbad <= 0.02 * len(synd)
    if nbad:
        assert first[bad].min() >= 8
        assert not (conv64_at[bad] < first[bad]).any()


@pytest.mark.parametrize("stem,p", [("[[90, 8, 10]]", 0.01), ("[[144, 12, 12]]", 0.01), ("[[144, 12, 12]]", 0.05)])
def test_zero_syndrome_shortcut_changes_no_output(stem, p):
    """Low error rates: a quarter to 40 % of the shots have an all-zero syndrome; with positive priors the reference returns the
    all-zero correction at its first check (decoding.py:69-73) and the warp kernels (float32, float64) retire such shots without running the
    iteration (its own instantiation, chosen when the priors imply >= 10 % error-free shots).  Hard decisions, flags, exit iterations
    and posterior LLRs must not change: with the shortcut disabled, chosen by the priors, and forced."""
    import os
    H, _ = load_code_file(stem)
    n = H.shape[1]
    rng = np.random.default_rng(21)
    err = (rng.random((6000, n)) < p).astype(np.uint8)
    synd = _synd(H, err)
    assert (synd.sum(axis=1) == 0).mean() > (0.15 if p == 0.01 else 0.0)
    code = _code(H, "min_sum")
    outs = []
    for env in ({"QLDPC_NO_ZERO_SHORTCUT": "1"}, {}, {"QLDPC_FORCE_ZERO_SHORTCUT": "1"}):
        for k in ("QLDPC_NO_ZERO_SHORTCUT", "QLDPC_FORCE_ZERO_SHORTCUT"):
            os.environ.pop(k, None)
        os.environ.update(env)
        try:
            a = code.bp_decode_batch(synd, _prior(p, n), "min_sum", 50, 0.8, 0.7, 25.0, precision=32)                 # LLRs of every shot
            b = code.bposd_decode_batch(synd, _prior(p, n), "min_sum", 50, 0.8, 0.7, 25.0, precision=32, osd_order=0)  # LLRs of failures only
            b = b + code.bposd_decode_batch(synd, _prior(p, n), "min_sum", 50, 0.8, 0.7, 25.0, precision=64, osd_order=0)   # the float64 kernel
        finally:
            for k in env:
                os.environ.pop(k, None)
        outs.append((a, b))
    for a, b in outs[1:]:
        for x, y in zip(a, outs[0][0]):
            assert np.array_equal(x, y)
        for x, y in zip(b, outs[0][1]):
            assert np.array_equal(x, y)
    zero = synd.sum(axis=1) == 0
    assert outs[1][0][1][zero].all() and not outs[1][0][0][zero].any() and (outs[1][0][3][zero] == 0).all()


def test_staged_kernel_matches_on_chip_kernel():
    """The HBM-staged instantiation runs the same arithmetic: forcing it on a small code must give
    bit-identical float64 results."""
    H, _ = load_code_file("[[72, 12, 6]]")
    n = H.shape[1]
    rng = np.random.default_rng(5)
    err = (rng.random((700, n)) < 0.06).astype(np.uint8)
    synd = _synd(H, err)
    code = _code(H, "min_sum")
    for prec in (64, 32):
        a = code.bp_decode_batch(synd, _prior(0.06, n), "min_sum", 40, 0.8, 0.7, 25.0, precision=prec)
        b = code.bp_decode_batch(synd, _prior(0.06, n), "min_sum", 40, 0.8, 0.7, 25.0, precision=prec, staged=True)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)
    # float64 sum-product: the on-chip path is the warp kernel with its own branch-free tanh / atanh (sp_math.cuh), the staged one calls
    # CUDA's -- last-ulp differences: identical decisions, flags and exit iterations on the shots that converge, LLRs to 1e-9 there
    a = code.bp_decode_batch(synd[:200], _prior(0.06, n), "sum_product", 30, precision=64)
    b = code.bp_decode_batch(synd[:200], _prior(0.06, n), "sum_product", 30, precision=64, staged=True)
    ok = a[1] & b[1]
    assert ok.sum() >= 0.7 * len(ok) and (a[1] == b[1]).mean() >= 0.99
    assert np.array_equal(a[0][ok], b[0][ok]) and np.array_equal(a[3][ok], b[3][ok])
    assert np.allclose(a[2][ok], b[2][ok], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("stem,p", [("[[72, 12, 6]]", 0.06), ("[[90, 8, 10]]", 0.05), ("[[108, 8, 10]]", 0.05),
                                    ("[[144, 12, 12]]", 0.05), ("[[288, 12, 18]]", 0.05)])
def test_tiled_kernel_bit_identical_to_thread_per_shot(stem, p):
    """The production T-lanes-per-shot kernel performs the same float32 operations in the same order as
    the thread-per-shot kernel: results must be bit-identical (hard, flag, exit iteration, LLRs), for
    T = 4 and 8, any refill threshold, ragged batch sizes."""
    H, _ = load_code_file(stem)
    n = H.shape[1]
    rng = np.random.default_rng(17)
    err = (rng.random((5003, n)) < p).astype(np.uint8)
    synd = _synd(H, err)
    code = _code(H, "min_sum")
    prior = _prior(p, n)
    kw = dict(variant="min_sum", max_iter=60, alpha=0.8, damping=0.7, clip=25.0, precision=32)
    ref = code.bp_decode_batch(synd, prior, staged=2, **kw)
    assert code.geometry(code.config(**kw))["kernel"] in ("tiled", "warp_per_shot")
    assert code.geometry(code.config(lanes_per_shot=8, **kw))["kernel"] == "tiled"
    assert code.geometry(code.config(staged=3, **kw))["kernel"] == "warp_per_shot"
    got = code.bp_decode_batch(synd, prior, staged=3, **kw)          # warp-per-shot kernel, messages in registers
    for x, y in zip(got, ref):
        assert np.array_equal(x, y), (stem, "warp_per_shot")
    cost = code.tune_warp_layout(0)                                  # the constructed labelling is bank-conflict free
    assert cost["floor"] == cost["current"] < cost["natural"], cost
    assert code.tune_warp_layout(-1)["current"] == cost["natural"]   # natural labelling of lanes: same results
    got = code.bp_decode_batch(synd, prior, staged=3, **kw)
    for x, y in zip(got, ref):
        assert np.array_equal(x, y), (stem, "warp_per_shot, natural labelling", cost)
    assert code.tune_warp_layout(8_000_000)["current"] == cost["floor"]
    for T in (4, 8):
        for rmin in (0, 1, 32 // T):
            got = code.bp_decode_batch(synd, prior, lanes_per_shot=T, refill_min=rmin, **kw)
            for x, y in zip(got, ref):
                assert np.array_equal(x, y), (stem, T, rmin)
    for B in (1, 7, 65):
        got = code.bp_decode_batch(synd[:B], prior, **kw)
        for x, y in zip(got, ref):
            assert np.array_equal(x, y[:B])
    # float64 min-sum (the bit-exact parity mode): warp-per-shot kernel by default
    k64 = dict(variant="min_sum", max_iter=60, alpha=0.8, damping=0.7, clip=25.0, precision=64)
    assert code.geometry(code.config(**k64))["kernel"] == "warp_per_shot"
    a = code.bp_decode_batch(synd[:2000], prior, staged=2, **k64)
    for extra in (dict(), dict(lanes_per_shot=8)):
        b = code.bp_decode_batch(synd[:2000], prior, **extra, **k64)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (stem, "float64", extra)
    # the other instantiations of the tiled kernel: float64 min-sum, sum-product (plain and damped) in both precisions
    for variant, prec, al, dm, cl in (("min_sum", 64, 0.8, 0.7, 25.0), ("sum_product", 64, 1.0, 1.0, 20.0),
                                      ("sum_product_sym", 64, 0.9, 0.8, 20.0), ("sum_product", 32, 1.0, 1.0, 20.0),
                                      ("sum_product_sym", 32, 0.9, 0.8, 20.0)):
        k2 = dict(variant=variant, max_iter=40, alpha=al, damping=dm, clip=cl, precision=prec)
        assert code.geometry(code.config(lanes_per_shot=8, **k2))["kernel"] == "tiled"
        a = code.bp_decode_batch(synd[:1500], prior, staged=2, **k2)
        b = code.bp_decode_batch(synd[:1500], prior, lanes_per_shot=8, **k2)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), (stem, variant, prec)
    # non-uniform prior and default parameters (alpha = damping = 1)
    pr = rng.uniform(1.5, 4.0, n)
    a = code.bp_decode_batch(synd[:800], pr, "min_sum", 30, precision=32, staged=2)
    for kw2 in (dict(), dict(staged=3), dict(lanes_per_shot=8), dict(lanes_per_shot=4)):
        b = code.bp_decode_batch(synd[:800], pr, "min_sum", 30, precision=32, **kw2)
        for x, y in zip(a, b):
            assert np.array_equal(x, y), kw2


@pytest.mark.parametrize("schedule", ["seq", "reference"])
def test_cta_per_shot_kernel_bit_identical_to_staged_kernel(schedule):
    """Space-time matrix 864 x 2592 (BASELINE config 4): the CTA-per-shot kernel (messages in registers, 12 warps per
    shot, rows of weight 7 and 8, columns of weight 1-3) performs the float32 operations of the HBM-staged
    thread-per-shot kernel in the same order: hard decisions, flags, exit iterations and LLRs must be bit-identical."""
    from qldpc_b200 import Code, graph
    from qldpc_b200.spaceTime import spaceTimeMatrix
    H, _ = load_code_file("[[144, 12, 12]]")
    Hst = spaceTimeMatrix(H, 12)
    m, n = Hst.shape
    sched = (graph.SEQ, graph.SEQ) if schedule == "seq" else graph.reference_schedule(Hst, "min_sum")
    code = Code(Hst, None, sched)
    rng = np.random.default_rng(23)
    err = (rng.random((301, n)) < 0.004).astype(np.uint8)
    synd = _synd(Hst, err)
    for (prior, kw) in ((_prior(0.004, n), dict(variant="min_sum", max_iter=40, alpha=0.8, damping=0.7, clip=25.0, precision=32)),
                        (rng.uniform(2.0, 7.0, n), dict(variant="min_sum", max_iter=25, precision=32)),            # alpha = damping = 1
                        (np.full(n, 30.0), dict(variant="min_sum", max_iter=10, alpha=0.9, damping=0.8, clip=20.0, precision=32))):  # prior > clip
        assert code.geometry(code.config(**kw))["kernel"] == "cta_per_shot"
        assert code.geometry(code.config(staged=1, **kw))["kernel"] == "hbm_staged"
        ref = code.bp_decode_batch(synd, prior, staged=1, **kw)
        got = code.bp_decode_batch(synd, prior, **kw)
        for x, y in zip(got, ref):
            assert np.array_equal(x, y), (schedule, kw)
    # float32 sum-product (psi domain) on the same mapping against the float64 staged kernel
    prior = _prior(0.004, n)
    for variant, kw in (("sum_product", {}), ("sum_product_sym", dict(alpha=0.9, damping=0.8, clip=20.0))):
        assert code.geometry(code.config(variant, 30, precision=32, **kw))["kernel"] == "cta_per_shot"
        ref = code.bp_decode_batch(synd, prior, variant, 30, precision=64, **kw)
        got = code.bp_decode_batch(synd, prior, variant, 30, precision=32, **kw)
        same = (got[1] == ref[1]) & (got[3] == ref[3]) & (got[0] == ref[0]).all(1)
        sel = ref[1] & same
        rel = np.abs(got[2][sel] - ref[2][sel]) / np.maximum(np.abs(ref[2][sel]), 1e-3)
        print(f"\n[space-time f32 {variant}] identical {same.mean():.4f}, q99 rel. LLR error {np.quantile(rel, 0.99):.2e}")
        assert same.mean() >= 0.97 and np.quantile(rel, 0.99) < 1e-4, (variant, same.mean(), np.quantile(rel, 0.99))


def test_spacetime_bp_staged(spacetime_golden):
    """BASELINE config 4 (scaled to 3 rounds of [[72,12,6]] so the golden file stays small): BP on the
    space-time matrix, message state staged in HBM."""
    from qldpc_b200.spaceTime import spaceTimeMatrix
    d = spacetime_golden
    H, _ = load_code_file("[[72, 12, 6]]")
    Hst = spaceTimeMatrix(H, 3)
    code = _code(Hst, "min_sum")
    prior = _prior(0.02, Hst.shape[1])
    hard, conv, llr, iters = code.bp_decode_batch(d["synd"], prior, "min_sum", 50, 0.8, 0.7, 25.0, precision=64)
    assert code.geometry(code.config("min_sum", 50, 0.8, 0.7, 25.0, 64))["staged"]
    assert np.array_equal(hard, d["ms_hard"]) and np.array_equal(conv, d["ms_conv"])
    assert np.array_equal(iters, d["ms_iter"]) and np.array_equal(llr, d["ms_llr"])
    f = np.nonzero(~conv)[0]
    sol = code.osd_decode_batch(d["synd"][f], llr[f], hard[f])
    assert np.array_equal(sol, d["ms_osd0"][f])


def test_edge_cases():
    H, _ = load_code_file("[[72, 12, 6]]")
    code = _code(H, "min_sum")
    n, m = 72, 36
    prior = _prior(0.05, n)
    # empty batch
    hard, conv, llr, iters = code.bp_decode_batch(np.zeros((0, m), np.uint8), prior)
    assert hard.shape == (0, n) and conv.shape == (0,)
    # zero syndrome converges at iteration 0 to the zero error
    hard, conv, llr, iters = code.bp_decode_batch(np.zeros((1, m), np.uint8), prior, precision=64)
    assert conv[0] and iters[0] == 0 and not hard.any()
    # ragged batch sizes, max_iter = 1
    rng = np.random.default_rng(3)
    g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))
    for B in (1, 31, 33, 257):
        err = (rng.random((B, n)) < 0.05).astype(np.uint8)
        synd = _synd(H, err)
        for mi in (1, 7):
            ref = O.decode_batch(g, synd, prior, O.MIN_SUM, mi, 0.8, 0.7, 25.0, osd_order=-1, want_llr=True)
            hard, conv, llr, iters = code.bp_decode_batch(synd, prior, "min_sum", mi, 0.8, 0.7, 25.0, precision=64)
            assert np.array_equal(hard.astype(np.uint8), ref["corr"]) and np.array_equal(conv, ref["converged"])
            assert np.array_equal(iters, ref["iters"]) and np.array_equal(llr, ref["llr"])
    # Steane (3 x 7, one syndrome word, row weight 4)
    Hs, _ = load_code_file("steane")
    cs = _code(Hs, "min_sum")
    gs = O.Graph(Hs)
    err = (rng.random((64, 7)) < 0.1).astype(np.uint8)
    ref = O.decode_batch(gs, _synd(Hs, err), _prior(0.1, 7), O.MIN_SUM, 20, 1.0, 1.0, 20.0, osd_order=0, want_llr=True)
    corr, conv, iters = cs.bposd_decode_batch(_synd(Hs, err), _prior(0.1, 7), "min_sum", 20, precision=64, osd_order=0)
    assert np.array_equal(corr, ref["corr"]) and np.array_equal(conv, ref["converged"])


# ----------------------------------------------------------------------------------------------
# OSD
# ----------------------------------------------------------------------------------------------
def test_osd_golden(osd_golden):
    d, meta = osd_golden
    for case in meta["cases"]:
        key = case["key"]
        H, _ = load_code_file(case["code"])
        code = _code(H, "loop")
        llr, hard = d[key + "_llr"], d[key + "_hard"]
        assert np.array_equal(code.osd_decode_batch(d[key + "_synd_c"], llr, hard), d[key + "_osd0_c"])
        assert np.array_equal(code.osd_decode_batch(d[key + "_synd_i"], llr, hard), d[key + "_osd0_i"])
        assert np.array_equal(code.osd_decode_batch(d[key + "_synd_c"], llr, hard, order=7), d[key + "_enh7_c"])
        for order, mc in case["sweeps"]:
            ref = d[f"{key}_enh_o{order}_mc{mc}"]
            k = len(ref)
            got = code.osd_decode_batch(d[key + "_synd_i"][:k], d[key + "_llr_tf"][:k], hard[:k], order=order,
                                        max_combinations=(mc or None))
            assert np.array_equal(got, ref), (case["code"], order, mc)


def test_osd_after_bp_with_ties(bp_golden):
    d, meta = bp_golden
    for case in meta["cases"]:
        key = case["key"]
        H, _ = load_code_file(case["code"], case["layout"])
        synd = _synd(H, d[key + "_errors"])
        code = _code(H, "loop")
        for pi in range(2):
            f = np.nonzero(~d[f"{key}_ms{pi}_conv"])[0]
            got = code.osd_decode_batch(synd[f], d[f"{key}_ms{pi}_llr"][f], d[f"{key}_ms{pi}_hard"][f])
            assert np.array_equal(got, d[f"{key}_ms{pi}_osd0"][f])


@pytest.mark.parametrize("hook", ["QLDPC_OSD_FORCE_BLOCK", "QLDPC_OSD_FORCE_ROWMAJOR", "QLDPC_OSD_FORCE_BLOCK+QLDPC_OSD_FORCE_ROWMAJOR"])
def test_osd_other_kernels_forced_on_small_codes(hook):
    """The production OSD-0 kernel on the reference's codes is the column-major warp kernel (test_osd_golden runs it).
    The block-per-shot kernels (large check matrices; column-major, and row-major behind it for inconsistent syndromes)
    and the row-major warp kernel (elimination record for OSD-w) must agree bit for bit with the golden vectors too:
    force each on the small codes in a fresh process (the hooks are read once per process)."""
    import os, subprocess, sys
    env = dict(os.environ, **{h: "1" for h in hook.split("+")})
    code = r"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
from conftest import load_code_file, GOLDEN
from qldpc_b200 import Code
d = np.load(os.path.join(GOLDEN, "osd_golden.npz")); meta = json.loads(str(d["meta"]))
for case in meta["cases"]:
    key = case["key"]; H, _ = load_code_file(case["code"]); code = Code(H)
    assert np.array_equal(code.osd_decode_batch(d[key + "_synd_c"], d[key + "_llr"], d[key + "_hard"]), d[key + "_osd0_c"]), case
    assert np.array_equal(code.osd_decode_batch(d[key + "_synd_i"], d[key + "_llr"], d[key + "_hard"]), d[key + "_osd0_i"]), case
b = np.load(os.path.join(GOLDEN, "bp_golden.npz")); bm = json.loads(str(b["meta"]))
for case in bm["cases"]:
    key = case["key"]; H, _ = load_code_file(case["code"], case["layout"]); code = Code(H)
    synd = ((b[key + "_errors"].astype(np.int64) @ np.asarray(H).T) % 2).astype(np.uint8)
    f = np.nonzero(~b[key + "_ms0_conv"])[0]
    assert np.array_equal(code.osd_decode_batch(synd[f], b[key + "_ms0_llr"][f], b[key + "_ms0_hard"][f]), b[key + "_ms0_osd0"][f]), case
print("BLOCK-OSD-OK")
"""
    from conftest import ROOT
    r = subprocess.run([sys.executable, "-c", code], env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert "BLOCK-OSD-OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


def test_spacetime_144x12_bp_osd_vs_oracle():
    """BASELINE config 4 at full size: H_st = spaceTimeMatrix(Hx_144, 12) (864 x 2592, 6840 edges), syndromes from
    spacetimeSyndrome, min-sum BP (HBM-staged float64 kernel) + OSD-0 (block-per-shot kernel): bit-exact vs the oracle."""
    from qldpc_b200.spaceTime import spaceTimeMatrix, spacetimeSyndrome
    H, _ = load_code_file("[[144, 12, 12]]")
    Hst = spaceTimeMatrix(H, 12)
    assert Hst.shape == (864, 2592)
    p = 0.001      # the sampler's first-block quirk (spaceTime.py:35) makes about half of these fail BP
    np.random.seed(12)
    synd = np.array([spacetimeSyndrome(H, p, 12)[1] for _ in range(48)], np.uint8)
    prior = _prior(p, Hst.shape[1])
    g = O.Graph(Hst, *O.auto_schedule(Hst, O.MIN_SUM))
    ref = O.decode_batch(g, synd, prior, O.MIN_SUM, 50, 0.8, 0.7, 25.0, osd_order=0, want_llr=True)
    code = _code(Hst, "min_sum")
    cfg = code.config("min_sum", 50, 0.8, 0.7, 25.0, 64)
    assert code.geometry(cfg)["kernel"] == "cta_staged"
    hard, conv, llr, iters = code.bp_decode_batch(synd, prior, "min_sum", 50, 0.8, 0.7, 25.0, precision=64)
    assert np.array_equal(conv, ref["converged"]) and np.array_equal(iters, ref["iters"]) and np.array_equal(llr, ref["llr"])
    assert (~conv).sum() >= 3, "want some BP failures to exercise OSD"
    corr, conv2, _ = code.bposd_decode_batch(synd, prior, "min_sum", 50, 0.8, 0.7, 25.0, precision=64, osd_order=0)
    assert np.array_equal(corr, ref["corr"])
    assert np.array_equal(_synd(Hst, corr), synd)          # H_st has full rank: every OSD solution satisfies its syndrome
    # standalone OSD call with float64 LLRs
    f = np.nonzero(~conv)[0]
    sol = code.osd_decode_batch(synd[f], llr[f], hard[f])
    assert np.array_equal(sol, ref["corr"][f])
    # float32 staged kernel: valid corrections
    corr32, _, _ = code.bposd_decode_batch(synd, prior, "min_sum", 50, 0.8, 0.7, 25.0, precision=32, osd_order=0)
    assert np.array_equal(_synd(Hst, corr32), synd)


@pytest.mark.parametrize("rounds,kind", [(12, "ties"), (12, "random"), (5, "ties"), (3, "random")])
def test_block_osd_rounds_of_32_candidates_vs_oracle(rounds, kind):
    """The block-per-shot OSD-0 kernel (rounds of 32 candidates: exclusive pivot rows, interaction matrix, folded row operations) on
    space-time matrices with LLRs that are NOT a BP output: a random order fills T in far more than a BP posterior does (many
    candidates without an exclusive bit, long reduction chains), and LLRs quantised to 8 levels put hundreds of columns in a tie
    (stable order: lower index first).  float32 and float64 keys; 12 rounds = 864 x 2592 (27-word kernel, register sort), 5 rounds =
    360 x 1080 (generic word count, register sort does not apply), 3 rounds = 216 x 648."""
    from qldpc_b200 import Code, graph
    from qldpc_b200.spaceTime import spaceTimeMatrix
    H, _ = load_code_file("[[144, 12, 12]]")
    Hst = spaceTimeMatrix(H, rounds).astype(np.int64)
    m, n = Hst.shape
    rng = np.random.default_rng(100 + rounds)
    B = 20
    err = (rng.random((B, n)) < 0.02).astype(np.uint8)
    synd = _synd(Hst, err)
    llr = rng.normal(2.0, 3.0, size=(B, n))
    if kind == "ties":
        llr = np.round(llr)                                                # ~8 distinct magnitudes, sign kept
        llr[llr == 0] = 1.0
    llr32 = llr.astype(np.float32).astype(np.float64)
    hard = (llr < 0).astype(np.uint8)
    g = O.Graph(Hst, *O.auto_schedule(Hst, O.MIN_SUM))
    code = Code(Hst, None, (graph.SEQ, graph.SEQ))
    for L in (llr, llr32):
        ref = np.stack([O.osd0(g, synd[b], L[b], hard[b]) for b in range(B)])
        got = code.osd_decode_batch(synd, L, hard)
        assert np.array_equal(got, ref)
        assert np.array_equal(_synd(Hst, got.astype(np.uint8)), synd)      # full row rank: every solution satisfies its syndrome
    # the float32-key instantiation (what the fused float32 BP -> OSD path launches), through the device-pointer ABI
    import torch
    from qldpc_b200 import _lib
    lib, dev, st = _lib.lib(), torch.device("cuda", 0), torch.cuda.current_stream().cuda_stream

    def pack(bits, words):
        u8 = torch.from_numpy(np.ascontiguousarray(bits, dtype=np.uint8)).to(dev)
        w = torch.zeros((B, words), dtype=torch.int32, device=dev)
        _lib.check(lib.qldpc_pack_bits_dev(u8.data_ptr(), w.data_ptr(), B, bits.shape[1], st))
        return w
    sw, hw = pack(synd, code.words_m), pack(hard, code.words_n)
    l32 = torch.from_numpy(llr32.astype(np.float32)).to(dev)
    out = torch.zeros((B, code.words_n), dtype=torch.int32, device=dev)
    valid = torch.zeros(B, dtype=torch.uint8, device=dev)
    _lib.check(lib.qldpc_osd_decode_dev(code.handle, None, None, B, sw.data_ptr(), l32.data_ptr(), 0, hw.data_ptr(), out.data_ptr(),
                                        valid.data_ptr(), st))
    torch.cuda.synchronize()
    ow = out.cpu().numpy().view(np.uint32)
    got32 = ((ow[:, np.arange(n) >> 5] >> (np.arange(n) & 31).astype(np.uint32)) & 1).astype(np.int64)
    assert np.array_equal(got32, ref) and bool(valid.all())


def test_osd_w_full_order7_vs_oracle():
    """41 225 candidates per shot on [[144,12,12]] (order 7, no cap): against the C oracle."""
    H, _ = load_code_file("[[144, 12, 12]]")
    m, n = H.shape
    rng = np.random.default_rng(21)
    g = O.Graph(H)
    llr = rng.normal(0, 4, (3, n))
    hard = (llr < 0).astype(np.uint8)
    synd = rng.integers(0, 2, (3, m)).astype(np.uint8)
    code = _code(H, "loop")
    got = code.osd_decode_batch(synd, llr, hard, order=7)
    for i in range(3):
        want, swept = O.osd_enhanced(g, synd[i], llr[i], hard[i], order=7)
        assert swept and np.array_equal(got[i], want)


# ----------------------------------------------------------------------------------------------
# fused pipeline, checks, sampler, Monte Carlo
# ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("stem,p", [("[[72, 12, 6]]", 0.05), ("[[144, 12, 12]]", 0.05), ("[[288, 12, 18]]", 0.06), ("[[90, 8, 10]]", 0.05)])
def test_bposd_pipeline_f64_bit_exact_vs_oracle(stem, p):
    H, d = load_code_file(stem)
    n = H.shape[1]
    rng = np.random.default_rng(2)
    B = 4000
    err = (rng.random((B, n)) < p).astype(np.uint8)
    synd = _synd(H, err)
    prior = _prior(p, n)
    g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))
    ref = O.decode_batch(g, synd, prior, O.MIN_SUM, 50, 0.8, 0.7, 25.0, osd_order=0)
    code = _code(H, "min_sum", d["Lx"], int(d["distance"]))
    corr, conv, iters = code.bposd_decode_batch(synd, prior, "min_sum", 50, 0.8, 0.7, 25.0, precision=64, osd_order=0)
    assert (~conv).sum() > 0
    assert np.array_equal(conv, ref["converged"]) and np.array_equal(iters, ref["iters"])
    assert np.array_equal(corr, ref["corr"])
    corr7, _, _ = code.bposd_decode_batch(synd, prior, "min_sum", 50, 0.8, 0.7, 25.0, precision=64, osd_order=7)
    assert np.array_equal(corr7, corr)
    # checks: bit-exact against the oracle's and against NumPy
    lg, va, wt = O.check_batch(g, d["Lx"], err, corr, synd)
    chk = code.check_batch(err, corr, synd, conv, iters)
    assert np.array_equal(chk["logical"], lg) and np.array_equal(chk["valid"], va) and np.array_equal(chk["weight"], wt)
    assert va.all()
    resid = corr ^ err
    assert np.array_equal(chk["logical"], ((d["Lx"].astype(np.int64) @ resid.T) % 2).any(0))
    c = chk["counters"]
    assert c["shots"] == B and c["bp_failed"] == int((~conv).sum()) and c["logical"] == int(lg.sum())
    assert c["logical_and_osd"] == int((lg & ~conv).sum()) and c["invalid"] == 0
    assert c["degenerate"] == int((va & ~lg & (corr != err).any(1)).sum())
    werr = err.sum(1)
    assert c["miscorrected"] == int((lg & (werr < int(d["distance"]) // 2)).sum())
    assert c["incorrectable"] == int((lg & (werr >= int(d["distance"]) // 2)).sum())
    assert c["iter_sum"] == int(iters.sum()) and c["residual_weight"] == int(wt.sum())
    # BP only
    corr_bp, conv_bp, _ = code.bposd_decode_batch(synd, prior, "min_sum", 50, 0.8, 0.7, 25.0, precision=64, osd_order=-1)
    ref_bp = O.decode_batch(g, synd, prior, O.MIN_SUM, 50, 0.8, 0.7, 25.0, osd_order=-1)
    assert np.array_equal(corr_bp, ref_bp["corr"]) and np.array_equal(conv_bp, conv)


def _philox_numpy(sid, blk, stream, seed):
    """NumPy replica of philox4x32_10 in csrc/misc_kernels.cuh."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    c = [np.asarray(sid & 0xffffffff, np.uint64), np.asarray(sid >> 32, np.uint64), np.asarray(blk, np.uint64), np.asarray(stream, np.uint64)]
    c = [np.broadcast_to(x, np.broadcast(*c).shape).astype(np.uint64) for x in c]
    k0, k1 = np.uint64(seed & 0xffffffff), np.uint64(seed >> 32)
    mask = np.uint64(0xffffffff)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return c


def test_syndrome_kernel_and_all_kernels_smoke():
    rng = np.random.default_rng(8)
    for stem in ("steane", "[[90, 8, 10]]", "[[288, 12, 18]]"):
        H, _ = load_code_file(stem)
        code = _code(H, "min_sum")
        err = (rng.random((1000, H.shape[1])) < 0.1).astype(np.uint8)
        assert np.array_equal(code.syndromes(err).astype(np.uint8), _synd(H, err))
    from qldpc_b200.spaceTime import spaceTimeMatrix
    Hst = spaceTimeMatrix(load_code_file("[[72, 12, 6]]")[0], 4)
    err = (rng.random((300, Hst.shape[1])) < 0.05).astype(np.uint8)
    assert np.array_equal(_code(Hst, "min_sum").syndromes(err).astype(np.uint8), _synd(Hst, err))
    import runpy
    from conftest import ROOT
    import os
    runpy.run_path(os.path.join(ROOT, "tools", "sanitize_smoke.py"), run_name="__main__")    # every kernel, small batches


def test_sampler_is_philox_keyed_by_global_shot_id():
    H, _ = load_code_file("[[144, 12, 12]]")
    code = _code(H, "min_sum")
    n, p, seed = 144, 0.05, 1234567890123
    err, synd = code.sample(p, 300, seed=seed, first_shot=1000)
    thr = np.uint64(int(p * 2 ** 32))
    sid = np.arange(1000, 1300, dtype=np.uint64)[:, None]
    j = np.arange(n)[None, :]
    r = _philox_numpy(sid, (j >> 2).astype(np.uint64), np.uint64(0), seed)
    words = np.stack(r, axis=-1)                                   # (B, n, 4)
    u = np.take_along_axis(words, (j & 3)[..., None].repeat(300, 0), axis=-1)[..., 0]
    want = (u < thr).astype(np.int8)
    assert np.array_equal(err, want)
    assert np.array_equal(synd.astype(np.uint8), _synd(H, err.astype(np.uint8)))
    assert abs(err.mean() - p) < 0.005
    # sharding invariance: two half ranges == one range
    a, _ = code.sample(p, 150, seed=seed, first_shot=1000)
    b, _ = code.sample(p, 150, seed=seed, first_shot=1150)
    assert np.array_equal(np.concatenate([a, b]), err)
    # draws = 2: XOR of two independent streams (paperResults.py:61-63)
    e2, s2 = code.sample(p, 300, seed=seed, first_shot=1000, draws=2)
    r1 = _philox_numpy(sid, (j >> 2).astype(np.uint64), np.uint64(1), seed)
    u1 = np.take_along_axis(np.stack(r1, axis=-1), (j & 3)[..., None].repeat(300, 0), axis=-1)[..., 0]
    assert np.array_equal(e2, want ^ (u1 < thr).astype(np.int8))


def test_mc_sweep_counters_match_host_pipeline_and_shard_invariance():
    H, d = load_code_file("[[144, 12, 12]]")
    code = _code(H, "min_sum", d["Lx"], int(d["distance"]))
    p, N, seed = 0.05, 20000, 7
    kw = dict(variant="min_sum", max_iter=50, alpha=0.8, damping=0.7, clip=25.0, precision=32, osd_order=0)
    whole = code.mc_sweep(p, N, seed=seed, **kw)
    assert whole["shots"] == N and whole["invalid"] == 0
    parts = [code.mc_sweep(p, N // 4, seed=seed, first_shot=i * (N // 4), **kw) for i in range(4)]
    for k in whole:
        assert whole[k] == sum(x[k] for x in parts), k
    err, synd = code.sample(p, N, seed=seed)
    corr, conv, iters = code.bposd_decode_batch(synd, _prior(p, 144), kw["variant"], 50, 0.8, 0.7, 25.0, precision=32, osd_order=0)
    chk = code.check_batch(err, corr, synd, conv, iters)
    assert chk["counters"] == whole
    # (the LER of this decoder against the reference's stored statistics: tests/test_gpu_ci.py)


def test_ler_matches_reference_stats_sum_product(reference_stats):
    """BPOSD.npz (sum-product BP50 + OSD-0, 10^4 shots): [[144,12,12]] LER 0.0499 at p ~ 0.0501.  Our LER on
    40 000 device-sampled shots must fall inside the reference's binomial 95 % CI (+ ours)."""
    st = reference_stats
    p = st["degeneracyCount_p"][7]
    want = st["BPOSD.npz"]["[[144, 12, 12]]"]["ler"][7]
    H, d = load_code_file("[[144, 12, 12]]")
    code = _code(H, "sum_product", d["Lx"], int(d["distance"]))
    N = 40000
    c = code.mc_sweep(p, N, seed=3, variant="sum_product", max_iter=50, precision=64, osd_order=0)
    ler = c["logical"] / N
    half = 1.96 * np.sqrt(want * (1 - want) / 10000) + 1.96 * np.sqrt(want * (1 - want) / N)
    print(f"\n[LER] sum-product BP50+OSD-0 [[144,12,12]] p={p:.4f}: ours {ler:.4f} vs reference {want:.4f} +- {half:.4f}")
    assert abs(ler - want) <= half
    assert c["invalid"] == 0


def test_kat_bp_npz_on_gpu(reference_stats):
    """notebooks/data/BP.npz [[72,12,6]] row (seed 0, sum-product BP50, BP only): the GPU float64 kernel on
    the reference's own RNG stream reproduces the stored counters."""
    st = reference_stats
    ps = np.logspace(-3.2, -1.3, 8)
    H, d = load_code_file("[[72, 12, 6]]")
    code = _code(H, "sum_product", d["Lx"], int(d["distance"]))
    want = st["BP.npz"]["[[72, 12, 6]]"]
    np.random.seed(0)
    n = 72
    dist = int(d["distance"])
    for pi, p in enumerate(ps):
        errors = (np.random.random((10000, n)) < p).astype(np.uint8)   # same stream as 10 000 draws of n
        synd = _synd(H, errors)
        corr, conv, iters = code.bposd_decode_batch(synd, _prior(p, n), "sum_product", 50, precision=64, osd_order=-1)
        chk = code.check_batch(errors, corr, synd, conv, iters)
        lg = chk["logical"]
        wt = errors.sum(1)
        got = dict(BPs_fault=int((~conv).sum()), degeneracies=int((conv & ~lg & (corr != errors).any(1)).sum()),
                   incorrectable=int((lg & (wt >= dist // 2)).sum()), BPs_miscorrected=int((lg & (wt < dist // 2)).sum()))
        got["ler"] = (got["BPs_fault"] + int(lg.sum())) / 10000
        for k, v in got.items():
            assert abs(v - want[k][pi]) <= (1 if k != "ler" else 2e-4), (pi, k, v, want[k][pi])


def test_full_size_properties():
    """BASELINE configs at scale, through size-independent properties: every BP+OSD correction satisfies its
    syndrome (invalid == 0), counters add up, 10^6 shots."""
    H, d = load_code_file("[[144, 12, 12]]")
    code = _code(H, "min_sum", d["Lx"], int(d["distance"]))
    N = 1_000_000
    c = code.mc_sweep(0.05, N, seed=99, variant="min_sum", max_iter=100, alpha=0.8, damping=0.7, clip=25.0, precision=32, osd_order=7)
    assert c["shots"] == N and c["invalid"] == 0
    assert c["logical"] == c["logical_and_osd"] + c["logical_and_bp_converged"]
    assert c["logical"] == c["miscorrected"] + c["incorrectable"]
    assert 0.02 < c["logical"] / N < 0.08
    H2, d2 = load_code_file("[[288, 12, 18]]")
    code2 = _code(H2, "min_sum", d2["Lx"], int(d2["distance"]))
    c2 = code2.mc_sweep(0.04, 200_000, seed=5, variant="min_sum", max_iter=50, alpha=0.8, damping=0.7, clip=25.0, precision=32, osd_order=-1)
    assert c2["shots"] == 200_000 and c2["invalid"] == c2["bp_failed"]   # BP-only: exactly the failures are invalid


# ----------------------------------------------------------------------------------------------
# arbitrary sparse H (circuit-level-DEM-like input) and error behaviour
# ----------------------------------------------------------------------------------------------
def _random_sparse_h(rng, m, n, max_col_w):
    H = np.zeros((m, n), np.uint8)
    for v in range(n):
        w = rng.integers(1, max_col_w + 1)
        H[rng.choice(m, size=w, replace=False), v] = 1
    return H


@pytest.mark.parametrize("m,n,cw,staged", [(60, 200, 5, 0), (60, 200, 5, 1), (170, 400, 4, 0), (24, 40, 7, 0)])
def test_arbitrary_sparse_h_nonuniform_priors_vs_oracle(m, n, cw, staged):
    """A detector-error-model-shaped input as in studies/studyComplete.py:80-99: irregular sparse H (scipy.sparse accepted),
    column weights up to 7, row weights well above 8, per-column priors from clipped probabilities (:88-89), an observables
    matrix.  float64 min-sum / sum-product BP + OSD-0 must equal the oracle bit for bit / to tolerance."""
    from scipy.sparse import csr_matrix
    from qldpc_b200 import Code
    rng = np.random.default_rng(m * 1000 + n)
    H = _random_sparse_h(rng, m, n, cw)
    probs = np.clip(rng.uniform(0.001, 0.08, n), 1e-15, 1 - 1e-15)
    prior = np.log((1 - probs) / probs)
    L = (rng.random((5, n)) < 0.1).astype(np.uint8)
    B = 600
    err = (rng.random((B, n)) < probs[None, :]).astype(np.uint8)
    synd = _synd(H, err)
    g = O.Graph(H)
    code = Code(csr_matrix(H), L)
    ref = O.decode_batch(g, synd, prior, O.MIN_SUM, 40, 0.8, 0.7, 25.0, osd_order=0, want_llr=True)
    hard, conv, llr, iters = code.bp_decode_batch(synd, prior, "min_sum", 40, 0.8, 0.7, 25.0, precision=64, staged=staged)
    assert np.array_equal(conv, ref["converged"]) and np.array_equal(iters, ref["iters"]) and np.array_equal(llr, ref["llr"])
    corr, conv2, _ = code.bposd_decode_batch(synd, prior, "min_sum", 40, 0.8, 0.7, 25.0, precision=64, osd_order=0, staged=staged)
    assert np.array_equal(corr, ref["corr"])
    assert (~conv).sum() > 0
    lg, va, wt = O.check_batch(g, L, err, corr, synd)
    chk = code.check_batch(err, corr, synd, conv, iters)
    if True:
        assert np.array_equal(chk["logical"], lg) and np.array_equal(chk["valid"], va) and np.array_equal(chk["weight"], wt)
    e2, s2 = code.sample(0.03, 500, seed=4)                     # device sampler on an arbitrary H (CSR syndromes when m > 160)
    assert np.array_equal(s2.astype(np.uint8), _synd(H, e2.astype(np.uint8)))
    cnt = code.mc_sweep(0.02, 2000, seed=1, variant="min_sum", max_iter=30, alpha=0.8, damping=0.7, clip=25.0, precision=32, osd_order=0)
    assert cnt["shots"] == 2000 and cnt["invalid"] <= cnt["bp_failed"]
    # sum-product on the same input: shots converging early agree to 1e-6
    refsp = O.decode_batch(g, synd, prior, O.SUM_PRODUCT, 30, osd_order=-1, want_llr=True)
    h2, c2, l2, i2 = code.bp_decode_batch(synd, prior, "sum_product", 30, precision=64, staged=staged)
    early = refsp["converged"] & (refsp["iters"] <= 10)
    assert early.sum() > B // 4
    assert np.array_equal(c2[early], refsp["converged"][early]) and np.array_equal(i2[early], refsp["iters"][early])
    np.testing.assert_allclose(l2[early], refsp["llr"][early], rtol=1e-6, atol=1e-9)


def test_error_behaviour():
    from qldpc_b200 import Code, QldpcError, _lib
    import ctypes
    H, _ = load_code_file("[[72, 12, 6]]")
    code = _code(H, "min_sum")
    prior = _prior(0.05, 72)
    with pytest.raises(ValueError):
        code.bp_decode_batch(np.zeros((4, 35), np.uint8), prior)                 # wrong number of checks
    with pytest.raises(ValueError):
        code.bp_decode_batch(np.zeros((4, 36), np.uint8), [1.0, 2.0])            # wrong prior length
    with pytest.raises(ValueError):
        code.osd_decode_batch(np.zeros((2, 36), np.uint8), np.zeros((2, 71)), np.zeros((2, 72), np.uint8))
    with pytest.raises(KeyError):
        code.bp_decode_batch(np.zeros((1, 36), np.uint8), prior, variant="nonsense")
    with pytest.raises(QldpcError):
        code.bp_decode_batch(np.zeros((1, 36), np.uint8), prior, max_iter=0)     # the C ABI validates its config
    with pytest.raises(QldpcError):
        code.bp_decode_batch(np.zeros((1, 36), np.uint8), prior, precision=16)
    with pytest.raises(ValueError):
        Code(np.zeros((3,)))                                                    # not a matrix
    L = _lib.lib()
    assert L.qldpc_code_create(0, 5, None, None, None, None, None, 0, None, ctypes.byref(ctypes.c_void_p())) != 0
    assert b"bad argument" in L.qldpc_last_error()
    # an all-zero prior (p = 0.5) and an all-ones syndrome are legal inputs
    hard, conv, llr, it = code.bp_decode_batch(np.ones((3, 36), np.uint8), np.zeros(72), "min_sum", 5, precision=64)
    assert hard.shape == (3, 72) and np.isfinite(llr).all()


# ----------------------------------------------------------------------------------------------
# OSD-w in the fused BP -> OSD paths (performOSD_enhanced with order > 0 on syndromes OUTSIDE the column space of H:
# measurement errors, the model the reference keeps one uncomment away, paperResults.py:66-68)
# ----------------------------------------------------------------------------------------------
def _noisy_case(stem, B, p, q, seed):
    H, d = load_code_file(stem)
    n = H.shape[1]
    rng = np.random.default_rng(seed)
    err = (rng.random((B, n)) < p).astype(np.uint8)
    synd = _synd(H, err) ^ (rng.random((B, H.shape[0])) < q).astype(np.uint8)      # syndrome + measurementError (mod 2)
    return H, d, err, synd


@pytest.mark.parametrize("stem,order,B", [("[[72, 12, 6]]", 2, 600), ("[[144, 12, 12]]", 3, 300), ("[[90, 8, 10]]", 7, 40),
                                          ("[[288, 12, 18]]", 1, 120)])
def test_fused_bposd_runs_the_osdw_sweep_on_inconsistent_syndromes(stem, order, B):
    H, d, err, synd = _noisy_case(stem, B, 0.04, 0.03, 11)
    n = H.shape[1]
    prior = _prior(0.04, n)
    g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))
    ref = O.decode_batch(g, synd, prior, O.MIN_SUM, 30, 0.8, 0.7, 25.0, osd_order=order)
    ref0 = O.decode_batch(g, synd, prior, O.MIN_SUM, 30, 0.8, 0.7, 25.0, osd_order=0)
    if stem == "[[72, 12, 6]]":      # here the sweep also changes the answer on some shots (the larger codes keep OSD-0: it
        assert (ref["corr"] != ref0["corr"]).any(1).sum() >= 3               # stays the best of the all-invalid candidates)
    code = _code(H, "min_sum", d["Lx"], int(d["distance"]))
    corr, conv, iters = code.bposd_decode_batch(synd, prior, "min_sum", 30, 0.8, 0.7, 25.0, precision=64, osd_order=order)
    assert np.array_equal(conv, ref["converged"]) and np.array_equal(iters, ref["iters"])
    assert np.array_equal(corr, ref["corr"])
    corr0, _, _ = code.bposd_decode_batch(synd, prior, "min_sum", 30, 0.8, 0.7, 25.0, precision=64, osd_order=0)
    assert np.array_equal(corr0, ref0["corr"])
    # float32 pipeline: the fused path must equal its own two halves (BP with LLR output, then the batched OSD call on them)
    hard, conv32, llr32, _ = code.bp_decode_batch(synd, prior, "min_sum", 30, 0.8, 0.7, 25.0, precision=32)
    f = np.nonzero(~conv32)[0]
    want = hard.astype(np.uint8)
    want[f] = code.osd_decode_batch(synd[f], llr32[f], hard[f], order=order).astype(np.uint8)
    corr32, c32, _ = code.bposd_decode_batch(synd, prior, "min_sum", 30, 0.8, 0.7, 25.0, precision=32, osd_order=order)
    assert np.array_equal(c32, conv32) and np.array_equal(corr32, want)


def test_mc_sweep_with_measurement_errors_equals_oracle_counters():
    """qldpc_mc_sweep_noisy == sample (device Philox, measurement flips included) -> oracle BP + OSD-w -> oracle checks."""
    H, d = load_code_file("[[72, 12, 6]]")
    n, dist = H.shape[1], int(d["distance"])
    code = _code(H, "min_sum", d["Lx"], dist)
    B, p, q, order = 3000, 0.03, 0.02, 2
    err, synd = code.sample(p, B, seed=5, meas_p=q)
    clean = _synd(H, err.view(np.uint8))
    flips = synd.view(np.uint8) ^ clean
    assert abs(flips.mean() - q) < 5 * np.sqrt(q / flips.size)              # measurement errors at the requested rate
    err0, synd0 = code.sample(p, B, seed=5)                                  # same data errors without them
    assert np.array_equal(err0, err) and np.array_equal(synd0.view(np.uint8), clean)
    prior = _prior(p, n)
    g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))
    ref = O.decode_batch(g, synd, prior, O.MIN_SUM, 40, 0.8, 0.7, 25.0, osd_order=order)
    lg, va, wt = O.check_batch(g, d["Lx"], err, ref["corr"], synd)
    c = code.mc_sweep(p, B, seed=5, variant="min_sum", max_iter=40, alpha=0.8, damping=0.7, clip=25.0, precision=64, osd_order=order,
                      meas_p=q)
    assert c["shots"] == B and c["bp_failed"] == int((~ref["converged"]).sum())
    assert c["logical"] == int(lg.sum()) and c["invalid"] == int((~va).sum()) and c["residual_weight"] == int(wt.sum())
    assert c["invalid"] > 0                                                  # inconsistent syndromes cannot be satisfied
    # shard invariance with measurement errors
    a = code.mc_sweep(p, 1000, seed=5, first_shot=0, variant="min_sum", max_iter=40, alpha=0.8, damping=0.7, clip=25.0, precision=64,
                      osd_order=order, meas_p=q)
    b = code.mc_sweep(p, 2000, seed=5, first_shot=1000, variant="min_sum", max_iter=40, alpha=0.8, damping=0.7, clip=25.0, precision=64,
                      osd_order=order, meas_p=q)
    assert all(a[k] + b[k] == c[k] for k in c)


def test_osdw_on_large_check_matrices_is_exact_or_refused():
    """More than 160 rows: no sweep kernel.  Full row rank (the space-time matrix): every syndrome is consistent, order 7 ==
    order 0 exactly as in the reference; rank-deficient: QLDPC_ERR_UNSUPPORTED, never a silent OSD-0."""
    from qldpc_b200 import Code, graph
    from qldpc_b200._lib import QldpcError
    from qldpc_b200.spaceTime import spaceTimeMatrix
    H144, _ = load_code_file("[[144, 12, 12]]")
    Hst = spaceTimeMatrix(H144, 3).astype(np.int64)                         # 216 x 648, has an identity block
    rng = np.random.default_rng(2)
    B = 24
    synd = (rng.random((B, Hst.shape[0])) < 0.3).astype(np.uint8)            # arbitrary syndromes: all consistent
    llr = rng.normal(size=(B, Hst.shape[1]))
    hard = (llr < 0).astype(np.uint8)
    cst = Code(Hst, None, (graph.SEQ, graph.SEQ))
    o0 = cst.osd_decode_batch(synd, llr, hard, order=0)
    assert np.array_equal(cst.osd_decode_batch(synd, llr, hard, order=7), o0)
    assert np.array_equal(_synd(Hst, o0.astype(np.uint8)), synd)
    Hbig = np.kron(np.eye(3, dtype=np.int64), np.asarray(H144))              # 216 x 432, rank 198 < 216
    cbig = Code(Hbig, None, (graph.SEQ, graph.SEQ))
    sb = (rng.random((4, 216)) < 0.3).astype(np.uint8)
    lb = rng.normal(size=(4, 432))
    assert cbig.osd_decode_batch(sb, lb, (lb < 0).astype(np.uint8), order=0).shape == (4, 432)
    with pytest.raises(QldpcError, match="rank-deficient"):
        cbig.osd_decode_batch(sb, lb, (lb < 0).astype(np.uint8), order=2)
    with pytest.raises(QldpcError, match="rank-deficient"):
        cbig.bposd_decode_batch(sb, [2.0] * 432, "min_sum", 5, precision=64, osd_order=2)


@pytest.mark.parametrize("stem", ["[[72, 12, 6]]", "[[144, 12, 12]]", "[[288, 12, 18]]"])
def test_osd_float64_keys_that_share_their_high_word(stem):
    """The OSD kernels rank float64 |LLR| keys by their high words and redo a shot with the full keys when two keys share the
    high word and differ in the low one: LLRs that differ only below 2^-20 relative (and exact ties) must still give the
    stable float64 order of the oracle."""
    H, d = load_code_file(stem)
    m, n = H.shape
    rng = np.random.default_rng(8)
    B = 48
    err = (rng.random((B, n)) < 0.06).astype(np.uint8)
    synd = _synd(H, err)
    base = rng.choice([0.75, 1.0, 3.5], size=(B, n))
    llr = base * (1.0 + rng.integers(0, 6, size=(B, n)) * 2.0 ** -44) * rng.choice([-1.0, 1.0], size=(B, n))
    llr[B // 2:] = (base * rng.choice([-1.0, 1.0], size=(B, n)))[B // 2:]          # second half: exact ties only
    hard = (llr < 0).astype(np.uint8)
    g = O.Graph(H)
    want = np.stack([O.osd0(g, synd[i], llr[i], hard[i]) for i in range(B)])
    code = _code(H, "loop")
    assert np.array_equal(code.osd_decode_batch(synd, llr, hard), want)


@pytest.mark.parametrize("schedule", ["seq", "reference"])
def test_cta_staged_kernel_vs_thread_per_shot_staged_kernel(schedule):
    """bp_stage_kernel (one CTA per shot, messages staged in global memory with 128-bit accesses, summaries on chip) against
    the thread-per-shot HBM-staged kernel on the 864 x 2592 space-time matrix: min-sum bit-identical in float32 and float64
    (hard decisions, flags, exit iterations, LLRs); the exact (tanh-domain) float64 sum-product agrees to the last ulps of
    the row product (taken in slot order instead of column order)."""
    from qldpc_b200 import Code, graph
    from qldpc_b200.spaceTime import spaceTimeMatrix
    H, _ = load_code_file("[[144, 12, 12]]")
    Hst = spaceTimeMatrix(H, 12)
    m, n = Hst.shape
    sched = (graph.SEQ, graph.SEQ) if schedule == "seq" else graph.reference_schedule(Hst, "min_sum")
    code = Code(Hst, None, sched)
    rng = np.random.default_rng(29)
    err = (rng.random((203, n)) < 0.004).astype(np.uint8)
    synd = _synd(Hst, err)
    for prec in (32, 64):
        for (prior, kw) in ((_prior(0.004, n), dict(variant="min_sum", max_iter=40, alpha=0.8, damping=0.7, clip=25.0)),
                            (rng.uniform(2.0, 7.0, n), dict(variant="min_sum", max_iter=25)),                                  # alpha = damping = 1
                            (np.full(n, 30.0), dict(variant="min_sum", max_iter=10, alpha=0.9, damping=0.8, clip=20.0))):       # prior > clip
            assert code.geometry(code.config(staged=5, precision=prec, **kw))["kernel"] == "cta_staged"
            ref = code.bp_decode_batch(synd, prior, staged=1, precision=prec, **kw)
            got = code.bp_decode_batch(synd, prior, staged=5, precision=prec, **kw)
            for x, y in zip(got, ref):
                assert np.array_equal(x, y), (schedule, prec, kw)
    assert code.geometry(code.config("min_sum", 40, 0.8, 0.7, 25.0, precision=64))["kernel"] == "cta_staged"     # float64: the default
    prior = _prior(0.004, n)
    for variant, kw in (("sum_product", {}), ("sum_product_sym", dict(alpha=0.9, damping=0.8, clip=20.0))):
        assert code.geometry(code.config(variant, 30, precision=64, **kw))["kernel"] == "cta_staged"
        ref = code.bp_decode_batch(synd, prior, variant, 30, precision=64, staged=1, **kw)
        got = code.bp_decode_batch(synd, prior, variant, 30, precision=64, **kw)
        same = (got[1] == ref[1]) & (got[3] == ref[3]) & (got[0] == ref[0]).all(1)
        sel = ref[1] & same
        rel = np.abs(got[2][sel] - ref[2][sel]) / np.maximum(np.abs(ref[2][sel]), 1e-3)
        print(f"\n[cta_staged f64 {variant}] identical {same.mean():.4f}, max rel. LLR error {rel.max():.2e}")
        assert same.mean() >= 0.99 and rel.max() < 1e-7, (variant, same.mean(), rel.max())      # (own tanh / atanh vs the math library)
        got32 = code.bp_decode_batch(synd, prior, variant, 30, precision=32, staged=5, **kw)       # float32, tanh domain
        assert (got32[1] == ref[1]).mean() > 0.9


def test_host_side_packing_of_the_uint8_interface_equals_device_side():
    """qldpc_bposd_decode_host with its uint8 rows packed by host threads (bit-packed rows cross PCIe, 37 bytes per [[144,12,12]]
    shot) against the same call with the rows packed by kernels (221 bytes): identical corrections, flags and iterations, over
    several chunks of the three-stream pipeline (QLDPC_HOST_CHUNK shrinks them), ragged batch sizes, a code whose row lengths are
    not multiples of 16 ([[90,8,10]]: plain stores instead of the non-temporal path), and the byte counters of the handle."""
    import os
    from qldpc_b200 import Code, graph
    os.environ["QLDPC_HOST_CHUNK"] = "4096"
    try:
        for stem, B in (("[[144, 12, 12]]", 40000 + 123), ("[[90, 8, 10]]", 9001), ("[[72, 12, 6]]", 1)):
            H, d = load_code_file(stem)
            m, n = H.shape
            code = Code(H, d["Lx"], graph.reference_schedule(H, "min_sum"))
            rng = np.random.default_rng(3)
            synd = _synd(H, (rng.random((B, n)) < 0.05).astype(np.uint8))
            res = {}
            for mode in (0, 1, 2):
                code.set_host_pack(mode)
                s0 = code.host_transfer_stats()
                res[mode] = code.bposd_decode_batch(synd, _prior(0.05, n), "min_sum", 40, 0.8, 0.7, 25.0, precision=32, osd_order=7)
                s1 = code.host_transfer_stats()
                assert s1["host_pack"] == mode
                wm, wn = (m + 31) // 32, (n + 31) // 32
                if mode < 2:
                    assert s1["h2d_bytes"] - s0["h2d_bytes"] == B * (m if mode == 0 else 4 * wm)
                    assert s1["d2h_bytes"] - s0["d2h_bytes"] == B * ((n if mode == 0 else 4 * wn) + 1 + 4)
                    assert (s1["chunks_host"] - s0["chunks_host"] > 0) == (mode == 1) and (s1["chunks_device"] - s0["chunks_device"] > 0) == (mode == 0)
                else:
                    assert B * 4 * wm <= s1["h2d_bytes"] - s0["h2d_bytes"] <= B * m
            for mode in (1, 2):
                for x, y in zip(res[0], res[mode]):
                    assert np.array_equal(x, y), (stem, mode)
            assert (_synd(H, res[1][0]) == synd).all()
            code.set_host_pack(-1)
            auto = code.bposd_decode_batch(synd, _prior(0.05, n), "min_sum", 40, 0.8, 0.7, 25.0, precision=32, osd_order=7)
            assert code.host_transfer_stats()["host_pack"] in (0, 1, 2) and np.array_equal(auto[0], res[0][0])
    finally:
        del os.environ["QLDPC_HOST_CHUNK"]
