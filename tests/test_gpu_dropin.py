"""Drop-in proof: the reference's own driver scripts, unmodified (baseline/_ref/, vendored by tools/vendor_reference.py),
executed twice -- on the reference's NumPy decoders and on the qldpc_b200 module swap (sys.modules aliasing of `decoding.*`,
`spaceTime`: qldpc_b200.compat) -- must print / store the same results.  Both runs consume the same np.random stream, the
min-sum / OSD kernels are bit-exact and the float64 sum-product agrees to ~1e-12, so the counters are compared exactly."""
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


def test_paper_results_py_trials_reduced(tmp_path):
    """paperResults.py:33-116 (sum-product maxIter 200 + performOSD, two draws per shot): trials and code list reduced."""
    edits = [(r"trials = 1000\b", "trials = 60"), (r'(?s)codes = \[.*?\]\n', 'codes = ["[[72, 12, 6]]", "[[90, 8, 10]]"]\n'),
             (r"code_labels = \[.*?\]", "code_labels = ['72', '90']")]
    ref = DH.run_script("paperResults.py", swap=False, edits=edits, workdir=str(tmp_path))
    got = DH.run_script("paperResults.py", swap=True, edits=edits, workdir=str(tmp_path))
    assert set(ref["results_OSD"]) == set(got["results_OSD"]) == {"[[72, 12, 6]]", "[[90, 8, 10]]"}
    for name in ref["results_OSD"]:
        for key in ("ler", "BPs_fault", "BPs_miscorrected", "incorrectable", "degeneracies"):
            assert ref["results_OSD"][name][key] == got["results_OSD"][name][key], (name, key)
    assert sum(sum(r["ler"]) for r in ref["results_OSD"].values()) > 0            # the reduced run still sees logical errors
    # the stored file has the reference's layout (loadResults.py reads results.item())
    saved = np.load(str(tmp_path / "data" / "BPOSD.npz"), allow_pickle=True)["results"].item()
    assert saved == got["results_OSD"]


def test_rework_main_py_trials_reduced(tmp_path):
    """rework/main.py:51-134 (4-tuple sum-product maxIter 100 + performOSD_enhanced order 7): trials and experiment list reduced."""
    edits = [(r"trials = 10000\b", "trials = 40"), (r"for exp in experiment:", "for exp in experiment[:2]:")]
    ref = DH.run_script("rework/main.py", swap=False, edits=edits, from_rework=True, workdir=str(tmp_path))
    got = DH.run_script("rework/main.py", swap=True, edits=edits, from_rework=True, workdir=str(tmp_path))
    assert list(ref["results"]) == list(got["results"]) == ["72", "90"]
    osd_rate = 0.0
    for name in ref["results"]:
        assert list(ref["results"][name]) == list(got["results"][name])
        for p, r in ref["results"][name].items():
            g = got["results"][name][p]
            assert set(r) == set(g)
            for key in r:
                assert np.array_equal(np.asarray(r[key]), np.asarray(g[key])), (name, p, key)
            osd_rate += r["osd"]
    assert osd_rate > 0                                                         # OSD-7 was invoked in the reduced run
