"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference (imported from
/root/reference, see tools/refimport.py) on seeded inputs.  Run in the build container:

    python tools/make_golden.py

The committed outputs are what pins the oracle (oracle/) and, through it and directly,
the CUDA path.  Inputs are drawn with np.random.default_rng(seed); every array needed
to replay a case is stored next to the reference's outputs.
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import refimport  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
CODES = os.path.join(ROOT, "qldpc_b200", "data", "codes")

BP_CASES = [  # (code file stem, layout, p, shots)
    ("[[72, 12, 6]]", "F", 0.05, 40),
    ("[[90, 8, 10]]", "F", 0.05, 24),
    ("[[108, 8, 10]]", "F", 0.05, 24),
    ("[[144, 12, 12]]", "F", 0.05, 40),
    ("[[144, 12, 12]]", "C", 0.06, 24),
    ("[[288, 12, 18]]", "F", 0.06, 24),
]
MINSUM_PARAMS = [(1.0, 1.0, 20.0), (0.8, 0.7, 25.0)]   # defaults; Alvarado-style (decoding.py:5, Alvarado.py:153)
SYM_PARAMS = (0.9, 0.8, 20.0)                           # decoding.py:131 defaults with alpha 0.9
MAX_ITER = 50


def load_code(stem, layout="F"):
    d = np.load(os.path.join(CODES, stem + ".npz"))
    H = d["Hx"]
    if layout == "C":
        H = np.ascontiguousarray(H)
    return H, d


def stable_ranks(llr):
    """Distinct magnitudes inducing the STABLE ascending order of |llr|."""
    order = np.argsort(np.abs(llr), kind="stable")
    r = np.empty(len(llr), np.float64)
    r[order] = np.arange(1, len(llr) + 1, dtype=np.float64)
    return r


def gen_bp(ref):
    out = {}
    meta = []
    for ci, (stem, layout, p, shots) in enumerate(BP_CASES):
        H, d = load_code(stem, layout)
        n = H.shape[1]
        rng = np.random.default_rng(1000 + ci)
        errors = (rng.random((shots, n)) < p).astype(np.uint8)
        synd = (errors.astype(np.int64) @ H.T) % 2
        prior = [np.log((1 - p) / p)] * n
        key = "c%d" % ci
        out[key + "_errors"] = errors
        meta.append(dict(key=key, code=stem, layout=layout, p=p, shots=shots, max_iter=MAX_ITER))
        for pi, (al, dm, cl) in enumerate(MINSUM_PARAMS):
            res = [ref.rework.performMinSum_Symmetric(H, synd[i], prior, maxIter=MAX_ITER, alpha=al, damping=dm, clip_llr=cl)
                   for i in range(shots)]
            out[f"{key}_ms{pi}_hard"] = np.array([r[0] for r in res], np.int8)
            out[f"{key}_ms{pi}_conv"] = np.array([r[1] for r in res], bool)
            out[f"{key}_ms{pi}_llr"] = np.array([r[2] for r in res], np.float64)
            out[f"{key}_ms{pi}_iter"] = np.array([r[3] for r in res], np.int32)
            # OSD on the BP failures, with tie-free rank LLRs (bit-exact contract, H1)
            osd = []
            for i in range(shots):
                if res[i][1]:
                    osd.append(res[i][0].astype(np.int64))
                else:
                    osd.append(ref.osd.performOSD(H, synd[i], stable_ranks(res[i][2]), res[i][0]))
            out[f"{key}_ms{pi}_osd0"] = np.array(osd, np.int8)
        res = [ref.bp.performBeliefPropagationFast(H, synd[i], prior, verbose=False, maxIter=MAX_ITER) for i in range(shots)]
        res4 = [ref.rework.performBeliefPropagationFast(H, synd[i], prior, maxIter=MAX_ITER) for i in range(shots)]
        for a, b in zip(res, res4):
            assert np.array_equal(a[0], b[0]) and a[1] == b[1] and np.array_equal(a[2], b[2])
        out[key + "_sp_hard"] = np.array([r[0] for r in res], np.int8)
        out[key + "_sp_conv"] = np.array([r[1] for r in res], bool)
        out[key + "_sp_llr"] = np.array([r[2] for r in res], np.float64)
        out[key + "_sp_iter"] = np.array([r[3] for r in res4], np.int32)
        al, dm, cl = SYM_PARAMS
        res = [ref.rework.performBeliefPropagation_Symmetric(H, synd[i], prior, maxIter=MAX_ITER, alpha=al, damping=dm, clip_llr=cl)
               for i in range(shots)]
        out[key + "_sym_hard"] = np.array([r[0] for r in res], np.int8)
        out[key + "_sym_conv"] = np.array([r[1] for r in res], bool)
        out[key + "_sym_llr"] = np.array([r[2] for r in res], np.float64)
        out[key + "_sym_iter"] = np.array([r[3] for r in res], np.int32)
        if layout == "C" or stem == "[[72, 12, 6]]":
            # loop version (beliefPropagation.py:6-85): same maths, sequential sums
            res = [ref.bp.performBeliefPropagation(H, synd[i], prior, verbose=False, maxIter=MAX_ITER) for i in range(min(shots, 12))]
            out[key + "_loop_hard"] = np.array([r[0] for r in res], np.int8)
            out[key + "_loop_conv"] = np.array([r[1] for r in res], bool)
            out[key + "_loop_llr"] = np.array([r[2] for r in res], np.float64)
        # alpha_estimation return paths on shot 0
        out[key + "_ms_alphaest"] = ref.rework.performMinSum_Symmetric(H, synd[0], prior, maxIter=1, alpha=0.8, damping=0.7,
                                                                      clip_llr=25.0, alpha_estimation=True)[2]
        out[key + "_sym_alphaest"] = ref.rework.performBeliefPropagation_Symmetric(H, synd[0], prior, maxIter=50, alpha=al, damping=dm,
                                                                                 clip_llr=cl, alpha_estimation=True)[2]
        print("bp golden", stem, layout, "min-sum fails:", [int((~out[f'{key}_ms{pi}_conv']).sum()) for pi in range(2)],
              "sp fails:", int((~out[key + '_sp_conv']).sum()))
    # non-uniform prior case (exercises the iteration-0 schedule)
    H, d = load_code("[[72, 12, 6]]", "F")
    n = H.shape[1]
    rng = np.random.default_rng(77)
    prior = rng.uniform(1.0, 4.0, n)
    errors = (rng.random((16, n)) < 0.06).astype(np.uint8)
    synd = (errors.astype(np.int64) @ H.T) % 2
    res = [ref.rework.performMinSum_Symmetric(H, synd[i], prior, maxIter=30, alpha=0.75, damping=0.7, clip_llr=25.0) for i in range(16)]
    out["nu_prior"] = prior
    out["nu_errors"] = errors
    out["nu_ms_hard"] = np.array([r[0] for r in res], np.int8)
    out["nu_ms_conv"] = np.array([r[1] for r in res], bool)
    out["nu_ms_llr"] = np.array([r[2] for r in res], np.float64)
    out["nu_ms_iter"] = np.array([r[3] for r in res], np.int32)
    res = [ref.bp.performBeliefPropagationFast(H, synd[i], prior, verbose=False, maxIter=30) for i in range(16)]
    out["nu_sp_hard"] = np.array([r[0] for r in res], np.int8)
    out["nu_sp_conv"] = np.array([r[1] for r in res], bool)
    out["nu_sp_llr"] = np.array([r[2] for r in res], np.float64)
    out["meta"] = np.array(json.dumps(dict(cases=meta, minsum_params=MINSUM_PARAMS, sym_params=SYM_PARAMS)))
    np.savez_compressed(os.path.join(OUT, "bp_golden.npz"), **out)


def gen_osd(ref):
    out = {}
    meta = []
    rng = np.random.default_rng(4242)
    for ci, (stem, nshots) in enumerate([("steane", 8), ("[[72, 12, 6]]", 24), ("[[90, 8, 10]]", 8), ("[[144, 12, 12]]", 24), ("[[288, 12, 18]]", 8)]):
        H, d = load_code(stem, "F")
        m, n = H.shape
        key = "o%d" % ci
        llr = rng.normal(0, 5, (nshots, n))                     # tie-free
        llr[:, : n // 8] = np.round(llr[:, : n // 8] * 4) / 4   # ... except a tied block, handled below
        hard = (llr < 0).astype(np.int8)
        err = (rng.random((nshots, n)) < 0.08).astype(np.int64)
        synd_c = (err @ H.T) % 2                                # consistent
        synd_i = rng.integers(0, 2, (nshots, m))                # (mostly) inconsistent
        ranks = np.array([stable_ranks(l) for l in llr])
        out[key + "_llr"] = llr
        out[key + "_hard"] = hard
        out[key + "_synd_c"] = synd_c.astype(np.uint8)
        out[key + "_synd_i"] = synd_i.astype(np.uint8)
        # OSD-0, consistent + inconsistent syndromes; reference fed rank-LLRs (stable contract)
        out[key + "_osd0_c"] = np.array([ref.osd.performOSD(H, synd_c[i], ranks[i], hard[i]) for i in range(nshots)], np.int8)
        out[key + "_osd0_i"] = np.array([ref.osd.performOSD(H, synd_i[i], ranks[i], hard[i]) for i in range(nshots)], np.int8)
        # OSD-w on consistent syndromes == OSD-0 (early return, OSD_enhanced.py:58-60)
        out[key + "_enh7_c"] = np.array([ref.osd_enh.performOSD_enhanced(H, synd_c[i], ranks[i], hard[i], order=7) for i in range(nshots)], np.int8)
        # OSD-w sweeps on inconsistent syndromes.  The metric uses |llr| VALUES, so feed the reference
        # tie-free LLRs directly: perturb the tied block with distinct tiny offsets.
        llr_tf = llr + np.arange(n) * 1e-7 * np.sign(llr + 1e-300)
        out[key + "_llr_tf"] = llr_tf
        sweeps = [(1, 0), (2, 0), (3, 60), (7, 200)] if n <= 144 else [(1, 0), (2, 40)]
        for order, mc in sweeps:
            k = nshots if (n <= 72 or mc or order == 1) else 6
            res = [ref.osd_enh.performOSD_enhanced(H, synd_i[i], llr_tf[i], hard[i], order=order, max_combinations=(mc or None)) for i in range(k)]
            out[f"{key}_enh_o{order}_mc{mc}"] = np.array(res, np.int8)
        meta.append(dict(key=key, code=stem, shots=nshots, sweeps=sweeps))
        print("osd golden", stem)
    out["meta"] = np.array(json.dumps(dict(cases=meta)))
    np.savez_compressed(os.path.join(OUT, "osd_golden.npz"), **out)


def gen_spacetime(ref):
    H, d = load_code("[[72, 12, 6]]", "F")
    Hst = ref.spacetime.spaceTimeMatrix(H, 3)
    np.random.seed(5)
    e, s = ref.spacetime.spacetimeSyndrome(H, 0.03, 3)
    # BP on the space-time matrix (studies/studyTT.py:49 calls the loop version; Fast gives the same triple)
    np.random.seed(6)
    cases = [ref.spacetime.spacetimeSyndrome(H, 0.02, 3) for _ in range(8)]
    p = 0.02
    prior = [np.log((1 - p) / p)] * Hst.shape[1]
    res = [ref.bp.performBeliefPropagationFast(Hst, c[1], prior, verbose=False, maxIter=50) for c in cases]
    ms = [ref.rework.performMinSum_Symmetric(Hst, c[1], prior, maxIter=50, alpha=0.8, damping=0.7, clip_llr=25.0) for c in cases]
    Hi = Hst.astype(np.int64)
    osd = [ref.osd.performOSD(Hi, c[1], stable_ranks(r[2]), r[0]) for c, r in zip(cases, ms)]
    np.savez_compressed(os.path.join(OUT, "spacetime_golden.npz"),
                        Hst_rows=np.nonzero(Hst)[0].astype(np.int32), Hst_cols=np.nonzero(Hst)[1].astype(np.int32),
                        Hst_shape=np.array(Hst.shape), Hst_is_c=np.array(Hst.flags["C_CONTIGUOUS"]),
                        seed5_error=e.astype(np.uint8), seed5_syndrome=s.astype(np.uint8),
                        synd=np.array([c[1] for c in cases], np.uint8),
                        sp_hard=np.array([r[0] for r in res], np.int8), sp_conv=np.array([r[1] for r in res], bool),
                        sp_llr=np.array([r[2] for r in res]),
                        ms_hard=np.array([r[0] for r in ms], np.int8), ms_conv=np.array([r[1] for r in ms], bool),
                        ms_llr=np.array([r[2] for r in ms]), ms_iter=np.array([r[3] for r in ms], np.int32),
                        ms_osd0=np.array(osd, np.int8))
    print("spacetime golden: BP fails", sum(not r[1] for r in ms))


def gen_stats(ref):
    """Stored result files of the reference that act as known-answer statistics."""
    st = {}
    bp = np.load(os.path.join(ref.root, "notebooks/data/BP.npz"), allow_pickle=True)["results"].item()
    bo = np.load(os.path.join(ref.root, "notebooks/data/BPOSD.npz"), allow_pickle=True)["results"].item()
    st["degeneracyCount_p"] = list(np.logspace(-3.2, -1.3, 8))
    st["BP.npz"] = {k: {kk: [float(x) for x in vv] for kk, vv in v.items()} for k, v in bp.items()}
    st["BPOSD.npz"] = {k: {kk: [float(x) for x in vv] for kk, vv in v.items()} for k, v in bo.items()}
    sim = np.load(os.path.join(ref.root, "rework/simulation_results.npz"), allow_pickle=True)["results"].item()
    st["simulation_results.npz"] = {
        code: {str(p): {k: float(v) for k, v in r.items() if np.isscalar(v) or np.ndim(v) == 0} for p, r in pr.items()}
        for code, pr in sim.items()}
    # rework/Alvarado.py:10-66 estimate_alpha_from_code, run unmodified (it imports `decoding` = rework/decoding.py
    # and matplotlib at module level, so only the function is executed, in a namespace with the names it uses)
    import ast
    src = open(os.path.join(ref.root, "rework/Alvarado.py")).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "estimate_alpha_from_code")
    from scipy.optimize import curve_fit
    ns = dict(np=np, curve_fit=curve_fit, performMinSum_Symmetric=ref.rework.performMinSum_Symmetric,
              performBeliefPropagation_Symmetric=ref.rework.performBeliefPropagation_Symmetric)
    exec(compile(ast.Module([fn], []), "Alvarado.py", "exec"), ns)
    import contextlib, io
    st["alpha_estimates"] = []
    for stem, p, trials, seed in [("[[72, 12, 6]]", 0.05, 600, 1), ("[[144, 12, 12]]", 0.05, 300, 2), ("[[72, 12, 6]]", 0.02, 600, 3)]:
        H, _ = load_code(stem, "F")
        np.random.seed(seed)
        with contextlib.redirect_stdout(io.StringIO()):
            a = ns["estimate_alpha_from_code"](H, trials=trials, error_rate=p, maxIter=1)
        st["alpha_estimates"].append(dict(code=stem, p=p, trials=trials, seed=seed, alpha=float(a)))
    with open(os.path.join(OUT, "reference_stats.json"), "w") as f:
        json.dump(st, f, indent=1)
    print("stats golden")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = refimport.load()
    which = sys.argv[1:] or ["bp", "osd", "spacetime", "stats"]
    if "bp" in which:
        gen_bp(ref)
    if "osd" in which:
        gen_osd(ref)
    if "spacetime" in which:
        gen_spacetime(ref)
    if "stats" in which:
        gen_stats(ref)
