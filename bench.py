#!/usr/bin/env python
"""bench.py -- BP+OSD decoded shots/s on the [[144,12,12]] gross code (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the decode hot path over one batch of synthetic syndromes:
min-sum BP (alpha 0.8, damping 0.7, clip 25, <= 100 iterations, float32) -> OSD-7 (performOSD_enhanced
semantics == OSD-0 on consistent syndromes) on the BP failures -> packed syndrome/logical checks and
LER counters, for 10^7 shots of i.i.d. bit-flip noise at p = 0.05 (code capacity), inputs already
resident in HBM.  `value` = shots of all ranks / max-over-ranks device time.  `e2e` = the same decode
through the host-pointer C-ABI call (qldpc_bposd_decode_host: uint8 syndromes in pinned host memory in,
uint8 corrections out), copies inside the timed region.

Multi-GPU (torchrun, one rank per GPU): every rank decodes its own shot-id range
[rank*B, (rank+1)*B) of the same Philox stream (weak scaling); the only collective is one NCCL
all-reduce of the counter vector.

`--impl reference`: the CPU arm -- the float64 C port of the reference's path (oracle/, OpenMP on
all host cores; the reference itself is pure Python and cannot travel to the GPU box) on a bounded
sample of the same workload.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CODE = "[[144, 12, 12]]"
P_ERR = 0.05
BP = dict(variant="min_sum", max_iter=100, alpha=0.8, damping=0.7, clip=25.0)
OSD_ORDER = 7
SEED = 0
METRIC = "BP+OSD decoded shots/sec, [[144,12,12]] gross code"
WORKLOAD = ("[[144,12,12]] Hx, code-capacity i.i.d. bit flips p=0.05, min-sum(alpha=0.8,damping=0.7,clip=25) BP<=100 iters "
            "+ OSD-7 (OSD_enhanced) on BP failures + syndrome/logical checks")


def load_code_arrays():
    d = np.load(os.path.join(ROOT, "qldpc_b200", "data", "codes", CODE + ".npz"))
    return d["Hx"], d["Lx"], int(d["distance"])


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_port_throughput(H, synd_u8, prior, budget_s=12.0):
    """The reference's path on the host cores: float64 C port (oracle/), all OpenMP threads, bounded sample."""
    from oracle import oracle as O
    g = O.Graph(H, O.SEQ, O.SEQ)
    # all host threads this process may use (torchrun exports OMP_NUM_THREADS=1: do not rely on the environment)
    try:
        cores = len(os.sched_getaffinity(0))
    except Exception:
        cores = os.cpu_count() or 1
    kw = dict(variant=O.MIN_SUM, max_iter=BP["max_iter"], alpha=BP["alpha"], damping=BP["damping"], clip=BP["clip"],
              osd_order=OSD_ORDER, nthreads=cores)
    probe = min(len(synd_u8), 20000)
    t0 = time.perf_counter()
    O.decode_batch(g, synd_u8[:probe], prior, **kw)
    rate = probe / (time.perf_counter() - t0)
    S = int(min(len(synd_u8), max(probe, rate * budget_s)))
    t0 = time.perf_counter()
    r = O.decode_batch(g, synd_u8[:S], prior, **kw)
    dt = time.perf_counter() - t0
    return S / dt, cores, S, dt, r


def run_reference(args, rank):
    if rank != 0:
        return
    H, Lx, dist = load_code_arrays()
    n = H.shape[1]
    prior = np.full(n, np.log((1 - P_ERR) / P_ERR))
    rng = np.random.default_rng(SEED)
    vals, sample = [], 0
    S = 2_000_000
    err = (rng.random((S, n)) < P_ERR).astype(np.uint8)
    synd = ((err.astype(np.int64) @ H.T) % 2).astype(np.uint8)
    for i in range(args.warmup + args.steps):
        v, cores, sample, dt, _ = cpu_port_throughput(H, synd, prior, budget_s=8.0)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": {"workload": WORKLOAD, "shots_per_step": sample},
            "cpu_baseline": {"value": value, "unit": "shots/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} shots of the same workload per step (float64 C port of beliefPropagation/min-sum + OSD, OpenMP)"},
            "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--shots", type=int, default=10_000_000, help="shots per step per GPU")
    ap.add_argument("--e2e-shots", type=int, default=0, help="shots per e2e step (default: same as --shots)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from qldpc_b200 import Code, _lib, graph

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    H, Lx, distance = load_code_arrays()
    m, n = H.shape
    code = Code(H, Lx, (graph.SEQ, graph.SEQ), distance)
    WM, WN = code.words_m, code.words_n
    cfg = Code.config(precision=32, **BP)
    prior = np.full(n, np.log((1 - P_ERR) / P_ERR))
    prior_p = prior.ctypes.data_as(ctypes.c_void_p)
    geom = code.geometry(cfg)
    if geom["kernel"] == "warp_per_shot":      # modelled scatter + gather wavefronts per shot-iteration of the lane labelling
        geom["warp_layout_wavefronts"] = code.tune_warp_layout(0)
    B = args.shots
    CH = 1 << 24          # shots per BP launch (LLR hand-off buffer: 4n bytes per shot)
    i32 = torch.int32
    err = torch.empty((B, WN), dtype=i32, device=dev)
    synd = torch.empty((B, WM), dtype=i32, device=dev)
    corr = torch.empty((B, WN), dtype=i32, device=dev)
    conv = torch.empty(B, dtype=torch.uint8, device=dev)
    iters = torch.empty(B, dtype=i32, device=dev)
    llr = torch.empty((min(B, CH), n), dtype=torch.float32, device=dev)
    fail = torch.zeros(min(B, CH) + 4, dtype=i32, device=dev)
    counters = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    iter_total = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(L.qldpc_sample_dev(code.handle, P_ERR, SEED, rank * B, 1, B, err.data_ptr(), synd.data_ptr(), stream), "sample")
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    kernel_ms = {"bp": 0.0, "osd": 0.0, "check": 0.0}
    launches = 0

    def step(timed):
        nonlocal launches
        marks = []
        for o in range(0, B, CH):
            b = min(CH, B - o)
            e0, e1, e2 = ev(), ev(), ev()
            e0.record()
            _lib.check(L.qldpc_bp_decode_dev(code.handle, ctypes.byref(cfg), prior_p, b, synd[o:].data_ptr(), corr[o:].data_ptr(),
                                             conv[o:].data_ptr(), iters[o:].data_ptr(), llr.data_ptr(), _lib.LLR_FAILED,
                                             fail[4:].data_ptr(), fail.data_ptr(), iter_total.data_ptr(), stream), "bp")
            e1.record()
            _lib.check(L.qldpc_osd_decode_dev(code.handle, fail[4:].data_ptr(), fail.data_ptr(), 0, synd[o:].data_ptr(), llr.data_ptr(),
                                              0, corr[o:].data_ptr(), corr[o:].data_ptr(), None, stream), "osd")
            e2.record()
            marks.append((e0, e1, e2))
            launches += 2
        e3, e4 = ev(), ev()
        e3.record()
        _lib.check(L.qldpc_check_dev(code.handle, B, err.data_ptr(), corr.data_ptr(), synd.data_ptr(), conv.data_ptr(),
                                     iters.data_ptr(), distance, None, None, counters.data_ptr(), stream), "check")
        e4.record()
        launches += 1
        return marks, (e3, e4)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step(False)
    barrier()
    counters.zero_()
    iter_total.zero_()
    launches = 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    t_start, t_end = ev(), ev()
    all_marks = []
    t_start.record()
    for _ in range(args.steps):
        all_marks.append(step(True))
    t_end.record()
    barrier()
    clocks = sampler.stop()
    total_ms = t_start.elapsed_time(t_end)
    for marks, (e3, e4) in all_marks:
        for e0, e1, e2 in marks:
            kernel_ms["bp"] += e0.elapsed_time(e1)
            kernel_ms["osd"] += e1.elapsed_time(e2)
        kernel_ms["check"] += e3.elapsed_time(e4)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    cnt = counters.clone()
    itot = iter_total.clone()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)      # the one data-path collective: LER counters
        dist.all_reduce(itot, op=dist.ReduceOp.SUM)
    total_ms = float(t.item())
    shots_total = B * args.steps * world
    value = shots_total / (total_ms * 1e-3)
    cnt = cnt.cpu().numpy()
    cd = dict(zip(_lib.COUNTER_NAMES, (int(x) for x in cnt)))

    # ---------------- e2e: host buffers through the C ABI ----------------
    Be = min(args.e2e_shots, B) if args.e2e_shots > 0 else B
    synd_u8_dev = torch.empty((Be, m), dtype=torch.uint8, device=dev)
    _lib.check(L.qldpc_unpack_bits_dev(synd.data_ptr(), synd_u8_dev.data_ptr(), Be, m, stream), "unpack")
    synd_h = torch.empty((Be, m), dtype=torch.uint8).pin_memory()
    synd_h.copy_(synd_u8_dev)
    corr_h = torch.empty((Be, n), dtype=torch.uint8).pin_memory()
    conv_h = torch.empty(Be, dtype=torch.uint8).pin_memory()
    iters_h = torch.empty(Be, dtype=torch.int32).pin_memory()
    torch.cuda.synchronize()

    def e2e_step():
        _lib.check(L.qldpc_bposd_decode_host(code.handle, ctypes.byref(cfg), prior_p, Be, synd_h.data_ptr(), OSD_ORDER,
                                             corr_h.data_ptr(), conv_h.data_ptr(), iters_h.data_ptr()), "bposd_host")

    e2e_steps = max(1, min(args.steps, 3))
    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = Be * e2e_steps * world / float(te.item())
    # same call with bit-packed host rows (not a reference dtype; reported next to the headline e2e)
    synd_hp = torch.empty((Be, WM), dtype=i32).pin_memory()
    synd_hp.copy_(synd[:Be])
    corr_hp = torch.empty((Be, WN), dtype=i32).pin_memory()
    torch.cuda.synchronize()

    def e2e_packed_step():
        _lib.check(L.qldpc_bposd_decode_host_packed(code.handle, ctypes.byref(cfg), prior_p, Be, synd_hp.data_ptr(), OSD_ORDER,
                                                    corr_hp.data_ptr(), conv_h.data_ptr(), iters_h.data_ptr()), "bposd_host_packed")
    e2e_packed_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_packed_step()
    torch.cuda.synchronize()
    tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    e2e_packed_value = Be * e2e_steps * world / float(tp.item())
    e2e_packed_matches = bool(torch.equal(corr_hp, corr[:Be].cpu()))
    # the host-path corrections must equal the device-path ones (same shots)
    corr_u8_dev = torch.empty((Be, n), dtype=torch.uint8, device=dev)
    _lib.check(L.qldpc_unpack_bits_dev(corr.data_ptr(), corr_u8_dev.data_ptr(), Be, n, stream), "unpack")
    torch.cuda.synchronize()
    e2e_matches = bool(torch.equal(corr_u8_dev.cpu(), corr_h))

    # ---------------- the other BP variants on the same workload (reported, not part of `value`) ----------------
    variants = {}
    if rank == 0:
        Bv = min(B, 1_000_000)
        for name, kw in (("min_sum_f64_bit_exact_vs_reference", dict(BP, precision=64)),
                         ("sum_product_f64", dict(variant="sum_product", max_iter=BP["max_iter"], precision=64)),
                         ("sum_product_f32", dict(variant="sum_product", max_iter=BP["max_iter"], precision=32))):
            cv = Code.config(**kw)
            run_v = lambda: _lib.check(L.qldpc_bposd_decode_dev(code.handle, ctypes.byref(cv), prior_p, Bv, synd.data_ptr(), OSD_ORDER,
                                                                corr.data_ptr(), conv.data_ptr(), iters.data_ptr(), None, stream), name)
            run_v()
            torch.cuda.synchronize()
            a, b_ = ev(), ev()
            a.record(); run_v(); b_.record()
            torch.cuda.synchronize()
            variants[name] = {"shots_per_s": Bv / (a.elapsed_time(b_) * 1e-3), "shots": Bv, "kernel": code.geometry(cv)["kernel"]}
        # the p-sweep of BASELINE configs[1] (rework/main.py:30 list + 0.01), same decoder, device-sampled syndromes
        sweep = {}
        Bs = min(B, 2_000_000)
        for pp in (0.01, 0.02, 0.03, 0.04, 0.05, 0.06, 0.1):
            _lib.check(L.qldpc_sample_dev(code.handle, pp, SEED + 1, 0, 1, Bs, err.data_ptr(), synd.data_ptr(), stream), "sample")
            pr = np.full(n, np.log((1 - pp) / pp))
            run_p = lambda: _lib.check(L.qldpc_bposd_decode_dev(code.handle, ctypes.byref(cfg), pr.ctypes.data_as(ctypes.c_void_p), Bs,
                                                                synd.data_ptr(), OSD_ORDER, corr.data_ptr(), conv.data_ptr(), iters.data_ptr(),
                                                                None, stream), "sweep")
            run_p()
            torch.cuda.synchronize()
            a, b_ = ev(), ev()
            a.record(); run_p(); b_.record()
            torch.cuda.synchronize()
            counters.zero_()
            _lib.check(L.qldpc_check_dev(code.handle, Bs, err.data_ptr(), corr.data_ptr(), synd.data_ptr(), conv.data_ptr(), iters.data_ptr(),
                                         distance, None, None, counters.data_ptr(), stream), "check")
            cs = dict(zip(_lib.COUNTER_NAMES, counters.cpu().tolist()))
            sweep[str(pp)] = {"shots_per_s": Bs / (a.elapsed_time(b_) * 1e-3), "ler": cs["logical"] / Bs,
                              "bp_failure_rate": cs["bp_failed"] / Bs, "invalid": cs["invalid"]}

    if rank == 0:
        A = 15 * code.E + 2 * n + m                       # lane-ops per shot-iteration
        iters_exec = int(itot.item()) / world             # per rank, over the timed steps
        bp_s = kernel_ms["bp"] * 1e-3
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_laneops = n_sm * 128 * sm_max * 1e6
        achieved = iters_exec * A / bp_s
        # packed syndrome in; packed correction + flag + iteration out; float32 posterior row for the BP failures (OSD input)
        hbm_bytes_per_shot = 4 * WM + 4 * WN + 1 + 4 + 4 * n * cd["bp_failed"] / max(1, cd["shots"])
        kname = {"tiled": "bp_tiled_kernel<T=%d,WM=%d,RW=6>" % (geom["lanes_per_shot"], WM),
                 "warp_per_shot": "bp_warp_kernel<CPL=%d,VPL=%d,RW=6>" % (WM, WN)}.get(geom["kernel"], "bp_decode_kernel<float,MIN_SUM>")
        roofline = {"bound": "alu", "kernel": kname,
                    "achieved": achieved / 1e12, "peak": peak_laneops / 1e12, "unit": "Tlane-op/s",
                    "frac": achieved / peak_laneops,
                    "peak_def": f"{n_sm} SMs x 128 lanes x {sm_max:.0f} MHz (max clock; median clock under this kernel {sm_mhz:.0f} MHz)",
                    "algorithmic_unit": f"A = 15E+2n+m = {A} lane-ops per shot-iteration; {iters_exec / (B * args.steps):.2f} iterations/shot executed",
                    "kernel_ms_per_step": kernel_ms["bp"] / args.steps,
                    "shot_iterations_per_s": iters_exec / bp_s,
                    "hbm": {"algorithmic_bytes_per_shot": hbm_bytes_per_shot,
                            "achieved_gbs": hbm_bytes_per_shot * B * args.steps / bp_s / 1e9,
                            "peak_gbs": peaks.get("hbm_gbs"), "note": "on-chip path: HBM traffic is negligible by design"},
                    "traffic": None}
        try:   # DRAM bytes per launch from the committed ncu capture of the same kernel (per shot x shots per launch)
            cap = json.load(open(os.path.join(ROOT, "profiles", "r1k_bp_ncu.json" if geom["kernel"] == "warp_per_shot" else "r1e_bp_ncu.json")))
            roofline["traffic"] = cap["dram_bytes_per_shot"] * min(B, CH)
            roofline["ncu_capture"] = {k: cap[k] for k in ("source", "issue_slots_busy_pct", "alu_pipe_pct", "lsu_pipe_pct",
                                                            "shared_wavefronts_pct_of_peak", "ipc_per_sm", "dram_bytes_per_shot")}
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "shots_per_step_per_gpu": B, "code": CODE, "p": P_ERR, "bp": BP, "osd_order": OSD_ORDER,
                           "l2": "per-step inputs+outputs (%.0f MB) exceed the 126 MB L2" % ((4 * WM + 8 * WN + 5) * B / 1e6),
                           "geometry": geom, "parallelism": f"shots sharded over {world} GPU(s), one all-reduce of counters"},
                "kernel_ms_per_step": {k: v / args.steps for k, v in kernel_ms.items()},
                "results": {"ler": cd["logical"] / max(1, cd["shots"]), "bp_failure_rate": cd["bp_failed"] / max(1, cd["shots"]),
                            "invalid": cd["invalid"], "mean_exit_iteration": cd["iter_sum"] / max(1, cd["shots"]), "shots": cd["shots"]},
                "roofline": roofline,
                "e2e": {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": Be * m, "d2h_bytes_per_step": Be * (n + 1 + 4),
                        "shots_per_step": Be, "steps": e2e_steps, "api": "qldpc_bposd_decode_host (uint8 in pinned host memory)",
                        "matches_device_path": e2e_matches,
                        "packed_host_rows": {"value": e2e_packed_value, "unit": "shots/s", "h2d_bytes_per_step": Be * 4 * WM,
                                             "d2h_bytes_per_step": Be * (4 * WN + 1 + 4), "api": "qldpc_bposd_decode_host_packed",
                                             "matches_device_path": e2e_packed_matches}},
                "other_variants_same_workload": variants, "p_sweep_configs1": sweep,
                "gpu_launches": launches, "clocks": clocks}
        if world == 1 and not args.no_cpu:
            synd_cpu = synd_h.numpy()
            v, cores, sample, dt, r = cpu_port_throughput(H, synd_cpu, prior)
            same = float((r["corr"] == corr_h[:sample].numpy()).all(1).mean())
            line["cpu_baseline"] = {"value": v, "unit": "shots/s", "cores": cores, "kind": "port",
                                    "sample": f"first {sample} shots of the same workload ({dt:.1f} s; float64 C port of min-sum BP + OSD, OpenMP)",
                                    "fraction_of_shots_with_identical_correction_vs_gpu_f32": same}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
