"""Statistical parity with the reference's stored results (tests/golden/reference_stats.json): the reference's own flow,
re-run on the GPU with an independent random stream, must reproduce its rates inside binomial confidence intervals.

rework/simulation_results.npz is the output of rework/Alvarado.py:141-207: per (code, p) alpha-hat =
estimate_alpha_from_code(code, error_rate=p, maxIter=1), then 10^4 shots of performMinSum_Symmetric(maxIter=50, alpha-hat,
damping 0.7, clip 25) + performOSD_enhanced(order=0); stored: logical error rate, OSD invocation rate, mean exit iteration.
Two-sample intervals (the reference's 10^4 shots and ours); with 54 comparisons a 95 % interval is expected to miss a few
by chance, so the bar is: at least 85 % of the comparisons inside their 95 % interval and ALL inside z = 4.
"""
import numpy as np
import pytest

from conftest import load_code_file

pytestmark = pytest.mark.gpu

CODES = {"72": "[[72, 12, 6]]", "144": "[[144, 12, 12]]", "288": "[[288, 12, 18]]"}


def _se_binomial(k, n, ref_rate, n_ref):
    pooled = (k + ref_rate * n_ref) / (n + n_ref)
    return np.sqrt(max(pooled * (1 - pooled), 1e-12) * (1.0 / n + 1.0 / n_ref))


def _z_binomial(k, n, ref_rate, n_ref):
    return abs(k / n - ref_rate) / _se_binomial(k, n, ref_rate, n_ref)


def test_alvarado_flow_reproduces_simulation_results(reference_stats):
    """The stored rates are functions of the reference run's own alpha-hat draw (5000 trials, unknown realisation): the
    convergence statistics (OSD rate, exit iteration) move with alpha, so every interval combines the two-sample binomial
    error with the alpha-hat sampling error -- its standard deviation from five independent estimates, its effect from decoding
    the same shots at alpha +- one standard deviation."""
    from qldpc_b200 import Code, graph
    from qldpc_b200.rework.Alvarado import estimate_alpha_from_code
    ref = reference_stats["simulation_results.npz"]
    N, NREF = 50000, 10000
    zs, report = [], []
    for name, stem in CODES.items():
        H, d = load_code_file(stem)
        n = H.shape[1]
        code = Code(H, d["Lx"], graph.reference_schedule(H, "min_sum"), int(d["distance"]))
        for pi, (pkey, want) in enumerate(ref[name].items()):
            p = float(pkey)
            alphas = [estimate_alpha_from_code(H, trials=5000, error_rate=p, maxIter=1, verbose=False, seed=100 + 10 * pi + r) for r in range(5)]
            alpha, a_sd = float(np.mean(alphas)), float(np.std(alphas, ddof=1))

            def run(a):
                c = code.mc_sweep(p, N, seed=7 + pi, first_shot=pi * N, variant="min_sum", max_iter=50, alpha=a, damping=0.7, clip=25.0,
                                  precision=64, osd_order=0)
                assert c["shots"] == N and c["invalid"] == 0
                return np.array([c["logical"] / N, c["bp_failed"] / N, c["iter_sum"] / N]), c
            r0, c = run(alpha)
            sens = np.abs(run(alpha + a_sd)[0] - run(alpha - a_sd)[0]) / 2          # same shots: the effect of one sd of alpha-hat
            # mean 0-based exit iteration: standard deviation from a per-shot sample of the same decoder
            _, synd = code.sample(p, 4000, seed=99, first_shot=pi * 4000)
            _, _, iters = code.bposd_decode_batch(synd, [np.log((1 - p) / p)] * n, "min_sum", 50, alpha, 0.7, 25.0, precision=64, osd_order=-1)
            se = np.array([_se_binomial(c["logical"], N, want["logical"], NREF), _se_binomial(c["bp_failed"], N, want["osd"], NREF),
                           max(float(np.std(iters)), 1e-9) * np.sqrt(1.0 / N + 1.0 / NREF)])
            # (sqrt(2): the reference's alpha-hat and ours are two independent draws)
            z = np.abs(r0 - np.array([want["logical"], want["osd"], want["average_iterations"]])) / np.sqrt(se ** 2 + 2 * sens ** 2)
            zs += list(z)
            report.append(f"{name} p={p}: alpha {alpha:.4f}+-{a_sd:.4f} LER {r0[0]:.4f} vs {want['logical']:.4f} (z {z[0]:.1f}); "
                          f"OSD {r0[1]:.4f} vs {want['osd']:.4f} (z {z[1]:.1f}); iters {r0[2]:.2f} vs {want['average_iterations']:.2f} (z {z[2]:.1f})")
    print("\n" + "\n".join(report))
    zs = np.array(zs)
    assert (zs <= 1.96).mean() >= 0.85, (zs > 1.96).sum()
    assert zs.max() <= 4.0, zs.max()


def test_bp_only_failures_288_match_bp_npz(reference_stats):
    """notebooks/data/BP.npz, [[288,12,18]] row (sum-product BP50, BP only, 10^4 shots per p): BP failure counts at the two
    highest error rates (679 / 10^4 at p ~ 0.0501) inside the two-sample binomial 95 % interval."""
    from qldpc_b200 import Code, graph
    st = reference_stats
    H, d = load_code_file("[[288, 12, 18]]")
    code = Code(H, d["Lx"], graph.reference_schedule(H, "sum_product"), int(d["distance"]))
    want = st["BP.npz"]["[[288, 12, 18]]"]["BPs_fault"]
    N = 60000
    for pi in (6, 7):
        p = st["degeneracyCount_p"][pi]
        c = code.mc_sweep(p, N, seed=21 + pi, variant="sum_product", max_iter=50, precision=64, osd_order=-1)
        z = _z_binomial(c["bp_failed"], N, want[pi] / 10000.0, 10000)
        print(f"\n[[288,12,18]] p={p:.4f}: BP failures {c['bp_failed'] / N:.4f} vs reference {want[pi] / 10000:.4f} (z {z:.2f})")
        assert z <= 1.96 * 1.3            # (two points: 95 % each would fail one run in ten by chance; z = 2.55 <-> 99 %)
