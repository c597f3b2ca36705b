"""Throughput of the other BASELINE.json configs (the headline config is bench.py's).  Device-resident inputs,
CUDA-event timing, one JSON object per line.

    python tools/bench_extras.py [--quick] > profiles/extras.jsonl
"""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from qldpc_b200 import Code, _lib, graph  # noqa: E402
from qldpc_b200.spaceTime import spaceTimeMatrix  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda", 0)
PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass


def load(name):
    d = np.load(os.path.join(ROOT, "qldpc_b200", "data", "codes", name + ".npz"))
    return d["Hx"], d["Lx"], int(d["distance"])


class Runner:
    def __init__(self, H, Lx=None, distance=0):
        self.code = Code(H, Lx, (graph.SEQ, graph.SEQ), distance)
        self.m, self.n = self.code.m, self.code.n
        self.distance = distance

    def run(self, p, B, cfg_kw, osd_order, synd_override=None, reps=2, prior_p=None):
        c = self.code
        WM, WN = c.words_m, c.words_n
        st = torch.cuda.current_stream().cuda_stream
        i32 = torch.int32
        err = torch.zeros((B, WN), dtype=i32, device=dev)
        synd = torch.zeros((B, WM), dtype=i32, device=dev)
        if synd_override is None:
            _lib.check(L.qldpc_sample_dev(c.handle, p, 1, 0, 1, B, err.data_ptr(), synd.data_ptr(), st))
        else:
            u8 = torch.from_numpy(synd_override).to(dev)
            _lib.check(L.qldpc_pack_bits_dev(u8.data_ptr(), synd.data_ptr(), B, self.m, st))
        corr = torch.empty((B, WN), dtype=i32, device=dev)
        conv = torch.empty(B, dtype=torch.uint8, device=dev)
        iters = torch.empty(B, dtype=i32, device=dev)
        itot = torch.zeros(1, dtype=torch.int64, device=dev)
        cnt = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
        cfg = Code.config(**cfg_kw)
        pp = p if prior_p is None else prior_p
        prior = np.full(self.n, np.log((1 - pp) / pp))

        def once():
            _lib.check(L.qldpc_bposd_decode_dev(c.handle, ctypes.byref(cfg), prior.ctypes.data_as(ctypes.c_void_p), B, synd.data_ptr(),
                                                osd_order, corr.data_ptr(), conv.data_ptr(), iters.data_ptr(), itot.data_ptr(), st))
        once()
        torch.cuda.synchronize()
        itot.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            once()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out = dict(shots=B, ms=ms, shots_per_s=B / ms * 1e3, iters_per_shot=itot.item() / reps / B,
                   bp_failure_rate=float(1.0 - conv.float().mean().item()), kernel=c.geometry(cfg)["kernel"],
                   lanes_per_shot=c.geometry(cfg)["lanes_per_shot"])
        if synd_override is None and c.k:
            _lib.check(L.qldpc_check_dev(c.handle, B, err.data_ptr(), corr.data_ptr(), synd.data_ptr(), conv.data_ptr(), iters.data_ptr(),
                                         self.distance, None, None, cnt.data_ptr(), st))
            torch.cuda.synchronize()
            cd = dict(zip(_lib.COUNTER_NAMES, cnt.cpu().tolist()))
            out.update(ler=cd["logical"] / B, invalid=cd["invalid"])
        out["shot_iterations_per_s"] = out["iters_per_shot"] * out["shots_per_s"]
        return out


def baseline_configs(quick=False, only=None, emit=None):
    """Runs BASELINE.json configs 1, 2 (p-sweep), 3, 5 and 4 with fixed shot counts; returns {label: result dict} (and calls
    emit(label, **result) as results arrive).  quick: a quarter of the shots (what bench.py embeds in its JSON line)."""
    q = 4 if quick else 1
    out = {}

    def put(config, **kw):
        if only is not None and config.split(":")[0] not in only:
            return
        out[config] = kw
        if emit:
            emit(config, **kw)
    want = lambda k: only is None or k in only
    ms_kw = dict(variant="min_sum", alpha=0.8, damping=0.7, clip=25.0, precision=32)
    ms64 = dict(ms_kw, precision=64)
    peak = 148 * 128 * PEAKS.get("sm_max_mhz", 1965.0) * 1e6

    def with_roof(r, res):
        """lane-op roofline fraction of the BP kernel alone (a BP-only run of the same shots)"""
        A = 15 * r.code.E + 2 * r.n + r.m
        if "bp_only" in res:
            res["bp_lane_op_frac"] = res["bp_only"]["shot_iterations_per_s"] * A / peak
        return res

    def both(r, p, B, cfg, osd, **kw):
        res = r.run(p, B, cfg, osd, **kw)
        if osd >= 0:
            b = r.run(p, B, cfg, -1, **kw)
            res["bp_only"] = dict(ms=b["ms"], shots_per_s=b["shots_per_s"], shot_iterations_per_s=b["shot_iterations_per_s"])
        else:
            res["bp_only"] = dict(ms=res["ms"], shots_per_s=res["shots_per_s"], shot_iterations_per_s=res["shot_iterations_per_s"])
        return with_roof(r, res)

    if want('1'):  # config 1: [[72,12,6]] p = 0.05, min-sum 50 iterations + OSD-0
        H, Lx, d = load("[[72, 12, 6]]")
        r = Runner(H, Lx, d)
        put("1: [[72,12,6]] p=0.05 min-sum BP50 + OSD-0, f32", **both(r, 0.05, 8_000_000 // q, dict(max_iter=50, **ms_kw), 0))
        put("1: [[72,12,6]] p=0.05 min-sum BP50 + OSD-0, f64 (bit-exact)", **both(r, 0.05, 4_000_000 // q, dict(max_iter=50, **ms64), 0))
        put("1: [[72,12,6]] p=0.05 min-sum defaults (alpha=1,damping=1,clip=20) BP50 + OSD-0, f32",
            **both(r, 0.05, 4_000_000 // q, dict(variant="min_sum", max_iter=50, precision=32), 0))

    if want('2'):  # config 2: [[144,12,12]] p-sweep, BP100 + OSD-7
        H, Lx, d = load("[[144, 12, 12]]")
        r = Runner(H, Lx, d)
        for p in (0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.1):
            put(f"2: [[144,12,12]] p={p} min-sum BP100 + OSD-7, f32", p=p, **both(r, p, 8_000_000 // q, dict(max_iter=100, **ms_kw), 7))
        for p in (0.01, 0.05, 0.1):
            put(f"2: [[144,12,12]] p={p} min-sum BP100 + OSD-7, f64 (bit-exact)", p=p, **both(r, p, 2_000_000 // q, dict(max_iter=100, **ms64), 7))
        put("2: [[144,12,12]] p=0.05 sum-product BP100 + OSD-7, f64", p=0.05,
            **both(r, 0.05, 400_000 // q, dict(variant="sum_product", max_iter=100, precision=64), 7, reps=1))
        put("2: [[144,12,12]] p=0.05 sum-product BP100 + OSD-7, f32 psi domain", p=0.05,
            **both(r, 0.05, 2_000_000 // q, dict(variant="sum_product", max_iter=100, precision=32), 7, reps=1))

    if want('3'):  # config 3: [[288,12,18]] BP only (BP failures and logical flags are separate counters)
        H, Lx, d = load("[[288, 12, 18]]")
        r = Runner(H, Lx, d)
        for p in (0.1, 0.06, 0.05, 0.04):
            put(f"3: [[288,12,18]] p={p} min-sum BP50, BP only, f32", p=p, **both(r, p, 4_000_000 // q, dict(max_iter=50, **ms_kw), -1))
        put("3: [[288,12,18]] p=0.05 min-sum BP50, BP only, f64 (bit-exact)", p=0.05, **both(r, 0.05, 1_000_000 // q, dict(max_iter=50, **ms64), -1))

    if want('5'):  # config 5: [[90,8,10]] / [[108,8,10]] p = 0.01, iteration budgets, min-sum vs sum-product
        for name in ("[[90, 8, 10]]", "[[108, 8, 10]]"):
            H, Lx, d = load(name)
            r = Runner(H, Lx, d)
            for mi in (10, 50, 90):
                put(f"5: {name} p=0.01 min-sum BP{mi} + OSD-0, f32", **both(r, 0.01, 8_000_000 // q, dict(max_iter=mi, **ms_kw), 0))
            put(f"5: {name} p=0.01 sum-product BP50 + OSD-0, f64", **both(r, 0.01, 1_000_000 // q, dict(variant="sum_product", max_iter=50, precision=64), 0, reps=1))
            put(f"5: {name} p=0.01 sum-product BP50 + OSD-0, f32 psi domain", **both(r, 0.01, 4_000_000 // q, dict(variant="sum_product", max_iter=50, precision=32), 0, reps=1))

    if want('4'):  # config 4: space-time [[144,12,12]] x 12 rounds (864 x 2592), BP50 + OSD-0
        H, Lx, d = load("[[144, 12, 12]]")
        Hst = spaceTimeMatrix(H, 12)
        r = Runner(Hst)
        E = r.code.E
        for p in (0.001, 0.005):
            B = 300_000 // q
            # the reference's own sampler semantics (first block = last round's syndrome), vectorised (spaceTime.py:20-43)
            rng = np.random.default_rng(4)
            m, n = H.shape
            err = (rng.random((B, n)) < p).astype(np.int64)
            s = (err @ H.T) % 2
            hist = []
            for _ in range(12):
                s = (s + (rng.random((B, m)) < p)) % 2
                hist.append(s)
            blocks = [hist[-1]] + [(hist[i] + hist[i - 1]) % 2 for i in range(1, 12)]
            synd = np.concatenate(blocks, axis=1).astype(np.uint8)
            res = r.run(p, B, dict(max_iter=50, **ms_kw), 0, synd_override=synd, reps=1)
            res_bp = r.run(p, B, dict(max_iter=50, **ms_kw), -1, synd_override=synd, reps=1)
            # the same BP with the message state staged in HBM (what BASELINE config 4 names): the crossover figure
            res_st = r.run(p, B, dict(max_iter=50, staged=1, **ms_kw), -1, synd_override=synd, reps=1)
            def staged_figures(rs, tsize=4):
                gbs = rs["shot_iterations_per_s"] * 3 * tsize * E / 1e9
                return dict(kernel=rs["kernel"], shots_per_s=rs["shots_per_s"], shot_iterations_per_s=rs["shot_iterations_per_s"],
                            algorithmic_bytes_per_shot_iteration=3 * tsize * E, hbm_algorithmic_gbs=gbs, hbm_frac=gbs / PEAKS.get("hbm_gbs", 6650.0))
            staged = staged_figures(res_st)
            # the CTA-per-shot kernel with the messages staged in global memory (bp_stage_kernel.cuh), float32 and float64
            cta_staged = staged_figures(r.run(p, B, dict(max_iter=50, staged=5, **ms_kw), -1, synd_override=synd, reps=1))
            cta_staged_f64 = staged_figures(r.run(p, B, dict(max_iter=50, **dict(ms_kw, precision=64)), -1, synd_override=synd, reps=1), 8)
            if res_bp["kernel"] == "cta_per_shot":
                A = 15 * E + 2 * r.n + r.m                       # lane-ops per shot-iteration
                roof = dict(bound="alu", algorithmic_lane_ops_per_shot_iteration=A, achieved=res_bp["shot_iterations_per_s"] * A / 1e12,
                            peak=peak / 1e12, unit="Tlane-op/s", frac=res_bp["shot_iterations_per_s"] * A / peak,
                            note="BP kernel alone (bp_only run); message state in registers / shared memory: HBM traffic negligible")
                label = "CTA-per-shot (on-chip)"
            else:
                bytes_per_iter = 12 * E
                gbs = res_bp["shot_iterations_per_s"] * bytes_per_iter / 1e9
                roof = dict(bound="hbm", algorithmic_bytes_per_shot_iteration=bytes_per_iter, achieved=gbs, peak=PEAKS.get("hbm_gbs", 6650.0),
                            unit="GB/s", frac=gbs / PEAKS.get("hbm_gbs", 6650.0), note="BP kernel alone (bp_only run)")
                label = "HBM-staged"
            put(f"4: space-time [[144,12,12]]x12 (864x2592, E={E}) p={p} min-sum BP50 + OSD-0, f32, {label}", p=p, **res,
                bp_only=dict(ms=res_bp["ms"], shots_per_s=res_bp["shots_per_s"], shot_iterations_per_s=res_bp["shot_iterations_per_s"]),
                roofline=roof, bp_only_hbm_staged=staged, bp_only_cta_staged=cta_staged, bp_only_cta_staged_f64=cta_staged_f64)
            # the bit-exact arithmetic on the same shots: float64 min-sum (bp_stage_kernel) + block OSD with float64 keys
            res64 = r.run(p, B // 2, dict(max_iter=50, **dict(ms_kw, precision=64)), 0, synd_override=synd[:B // 2], reps=1)
            put(f"4: space-time [[144,12,12]]x12 (864x2592, E={E}) p={p} min-sum BP50 + OSD-0, f64 (bit-exact), CTA-per-shot (staged)", p=p, **res64)
            if p == 0.001:      # the reference decodes these matrices with sum-product (studies/studyTT.py:49)
                sp32 = r.run(p, B, dict(variant="sum_product", max_iter=50, precision=32), 0, synd_override=synd, reps=1)
                sp64 = r.run(p, B // 4, dict(variant="sum_product", max_iter=50, precision=64), 0, synd_override=synd[:B // 4], reps=1)
                sp64b = r.run(p, B // 4, dict(variant="sum_product", max_iter=50, precision=64), -1, synd_override=synd[:B // 4], reps=1)
                sp64old = r.run(p, B // 16, dict(variant="sum_product", max_iter=50, precision=64, staged=1), -1, synd_override=synd[:B // 16], reps=1)
                put(f"4: space-time [[144,12,12]]x12 (864x2592) p={p} sum-product BP50 + OSD-0, f32 psi domain (CTA-per-shot)", p=p, **sp32,
                    float64=dict(shots=sp64["shots"], shots_per_s=sp64["shots_per_s"], kernel=sp64["kernel"],
                                 bp_only_shots_per_s=sp64b["shots_per_s"], bp_only_shot_iterations_per_s=sp64b["shot_iterations_per_s"],
                                 thread_per_shot_staged_bp_only_shot_iterations_per_s=sp64old["shot_iterations_per_s"]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="", help="comma-separated config numbers, e.g. 4")
    args = ap.parse_args()
    baseline_configs(args.quick, set(args.only.split(",")) if args.only else None,
                     emit=lambda config, **kw: print(json.dumps(dict(config=config, **kw)), flush=True))


if __name__ == "__main__":
    main()
