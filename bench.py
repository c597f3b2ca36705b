#!/usr/bin/env python
"""bench.py -- BP+OSD decoded shots/s on the [[144,12,12]] gross code (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One "step" = one pass of the decode hot path over one batch of synthetic syndromes: min-sum BP (alpha 0.8, damping 0.7,
clip 25, <= 100 iterations) -> OSD-0 on the BP failures -> OSD-7 sweep (performOSD_enhanced) on the shots whose OSD-0
solution misses the syndrome (none for syndromes of the form e*H^T, as in the reference) -> packed syndrome / logical checks
and LER counters, for 10^7 shots of i.i.d. bit-flip noise at p = 0.05 (code capacity), inputs resident in HBM.  The one
data-path collective -- the all-reduce of the counter vector -- is inside the timed region.

Two arms of the same workload, each with its own value, roofline and e2e:
  * the JSON line's `value`: float32 production kernels (north_star: "fp32 min-sum"); the fraction of shots whose
    correction equals the float64 CPU port's is reported next to it (non-converging trajectories diverge: DESIGN.md 4);
  * `bit_exact_f64`: the float64 kernels, bit-identical to the reference's arithmetic (checked against the CPU port
    inside this run).
`e2e` = the same decode through the host-pointer C-ABI call (qldpc_bposd_decode_host: uint8 syndromes in pinned host
memory in, uint8 corrections out), copies inside the timed region; `e2e.mc_sweep` = the reference's actual Monte-Carlo
use through qldpc_mc_sweep (sample -> decode -> check on the device, counters back).
`configs`: the other BASELINE.json configs (1, 3, 4, 5) with fixed shot counts and the same CUDA-event timing.
`sustained`: config 3 ([[288,12,18]] BP-only, 10^9 shots sharded over the ranks, counters all-reduced) with its clocks.

Multi-GPU (torchrun, one rank per GPU): every rank decodes its own shot-id range [rank*B, (rank+1)*B) of the same Philox
stream (weak scaling).

`--impl reference`: the CPU arm -- the float64 C port of the reference's path (oracle/, OpenMP on all host cores) on a
bounded sample of the same workload; the unmodified Python functions (baseline/_ref/, vendored by
tools/vendor_reference.py) are timed beside it on a smaller sample.
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
            "+ OSD-7 (OSD_enhanced: OSD-0, then the order-7 sweep where OSD-0 misses the syndrome) on BP failures "
            "+ syndrome/logical checks")


def load_code_arrays(name=CODE):
    d = np.load(os.path.join(ROOT, "qldpc_b200", "data", "codes", name + ".npz"))
    return d["Hx"], d["Lx"], int(d["distance"])


def base_config(shots):
    return {"workload": WORKLOAD, "shots_per_step_per_gpu": shots, "code": CODE, "p": P_ERR, "bp": BP, "osd_order": OSD_ORDER}


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
        pw = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [nm for i, nm in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        busy = [x for x in sm if x > 0]
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": reasons, "samples": len(sm)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))          # (torchrun exports OMP_NUM_THREADS=1: do not rely on the environment)
    except Exception:
        return os.cpu_count() or 1


def cpu_port_throughput(H, synd_u8, prior, budget_s=12.0):
    """The reference's path on the host cores: float64 C port (oracle/), all OpenMP threads, bounded sample."""
    from oracle import oracle as O
    g = O.Graph(H, *O.auto_schedule(H, O.MIN_SUM))        # the summation order NumPy uses for this (Fortran-ordered) Hx
    cores = host_threads()
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


def reference_python_throughput(H, synd_u8, prior, budget_s=8.0, check=None):
    """The UNMODIFIED reference functions (rework/decoding.py:5 performMinSum_Symmetric + :193 performOSD_enhanced, vendored
    under baseline/_ref/) on one host core -- the reference is single-threaded -- on a small sample of the same workload.
    Returns None when baseline/_ref is absent."""
    path = os.path.join(ROOT, "baseline", "_ref", "rework", "decoding.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    spec = importlib.util.spec_from_file_location("_reference_rework_decoding", path)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    Hd = np.asarray(H)
    pr = list(prior)
    n_done, same = 0, 0
    limit = len(synd_u8) if check is None else min(len(synd_u8), len(check))
    t0 = time.perf_counter()
    while n_done < limit and (time.perf_counter() - t0 < budget_s or n_done < 8):
        s = synd_u8[n_done].astype(np.int64)
        det, ok, llr, _ = ref.performMinSum_Symmetric(Hd, s, pr, maxIter=BP["max_iter"], alpha=BP["alpha"], damping=BP["damping"],
                                                      clip_llr=BP["clip"])
        if not ok:
            det = ref.performOSD_enhanced(Hd, s, llr, det, order=OSD_ORDER)
        if check is not None:
            same += int(np.array_equal(np.asarray(det).astype(np.uint8), check[n_done]))
        n_done += 1
    dt = time.perf_counter() - t0
    out = {"value": n_done / dt, "unit": "shots/s", "cores": 1, "kind": "reference_python",
           "sample": f"first {n_done} shots of the same workload ({dt:.1f} s; unmodified rework/decoding.py performMinSum_Symmetric + performOSD_enhanced(order=7), one process)"}
    # (corrections differ from the stable-sort contract only where np.argsort's default, unstable sort orders tied |LLR| values
    #  differently -- both are valid OSD orderings; tests/test_gpu_parity.py::test_osd_after_bp_with_ties feeds it the stable ranks)
    if check is not None:
        out["shots_with_correction_identical_to_gpu_f64"] = f"{same}/{n_done}"
    return out


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
    cpu = {"value": value, "unit": "shots/s", "cores": cores, "kind": "port",
           "sample": f"{sample} shots of the same workload per step (float64 C port of min-sum BP + OSD, OpenMP)"}
    rp = reference_python_throughput(H, synd, prior, budget_s=8.0)
    if rp:
        cpu["reference_python"] = rp
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": base_config(args.shots),
            "cpu_baseline": cpu,
            "e2e": {"value": value, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--shots", type=int, default=10_000_000, help="shots per step per GPU")
    ap.add_argument("--e2e-shots", type=int, default=0, help="shots per e2e step (default: same as --shots)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs / sustained / variant blocks")
    ap.add_argument("--sustained-shots", type=int, default=1_000_000_000, help="config 3 shots over all ranks")
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
    code = Code(H, Lx, graph.reference_schedule(H, "min_sum"), distance)     # the addition order NumPy uses on this Hx
    WM, WN = code.words_m, code.words_n
    prior = np.full(n, np.log((1 - P_ERR) / P_ERR))
    prior_p = prior.ctypes.data_as(ctypes.c_void_p)
    B = args.shots
    CH = 1 << 24          # shots per BP launch (LLR hand-off buffer: 4n / 8n bytes per shot)
    i32 = torch.int32
    err = torch.empty((B, WN), dtype=i32, device=dev)
    synd = torch.empty((B, WM), dtype=i32, device=dev)
    corr = torch.empty((B, WN), dtype=i32, device=dev)
    conv = torch.empty(B, dtype=torch.uint8, device=dev)
    iters = torch.empty(B, dtype=i32, device=dev)
    valid = torch.empty(B, dtype=torch.uint8, device=dev)
    llr = torch.empty((min(B, CH), n), dtype=torch.float64, device=dev)       # (float32 arm uses the first half)
    fail = torch.zeros(min(B, CH) + 4, dtype=i32, device=dev)
    counters = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    iter_total = torch.zeros(1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    _lib.check(L.qldpc_sample_dev(code.handle, P_ERR, SEED, rank * B, 1, B, err.data_ptr(), synd.data_ptr(), stream), "sample")
    torch.cuda.synchronize()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = peaks.get("sm_max_mhz", 1965.0)
    n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
    peak_laneops = n_sm * 128 * sm_max * 1e6

    ev = lambda: torch.cuda.Event(enable_timing=True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # host buffers of the e2e legs (shared by both arms)
    Be = min(args.e2e_shots, B) if args.e2e_shots > 0 else B
    synd_u8_dev = torch.empty((Be, m), dtype=torch.uint8, device=dev)
    _lib.check(L.qldpc_unpack_bits_dev(synd.data_ptr(), synd_u8_dev.data_ptr(), Be, m, stream), "unpack")
    synd_h = torch.empty((Be, m), dtype=torch.uint8).pin_memory()
    synd_h.copy_(synd_u8_dev)
    del synd_u8_dev
    corr_h = torch.empty((Be, n), dtype=torch.uint8).pin_memory()
    conv_h = torch.empty(Be, dtype=torch.uint8).pin_memory()
    iters_h = torch.empty(Be, dtype=torch.int32).pin_memory()
    synd_hp = torch.empty((Be, WM), dtype=i32).pin_memory()
    synd_hp.copy_(synd[:Be])
    corr_hp = torch.empty((Be, WN), dtype=i32).pin_memory()
    torch.cuda.synchronize()

    def run_arm(precision):
        """K timed steps of the device-resident hot path + the e2e legs, for one arithmetic.  Returns a dict."""
        cfg = Code.config(precision=precision, **BP)
        geom = code.geometry(cfg)
        if geom["kernel"] == "warp_per_shot":
            geom["warp_layout_wavefronts"] = code.tune_warp_layout(0)
        f64 = 1 if precision == 64 else 0
        kernel_ms = {"bp": 0.0, "osd0": 0.0, "osdw_sweep_stage": 0.0, "check": 0.0}
        launches = 0

        def step():
            nonlocal launches
            marks = []
            for o in range(0, B, CH):
                b = min(CH, B - o)
                e0, e1, e2, e3 = ev(), ev(), ev(), ev()
                e0.record()
                _lib.check(L.qldpc_bp_decode_dev(code.handle, ctypes.byref(cfg), prior_p, b, synd[o:].data_ptr(), corr[o:].data_ptr(),
                                                 conv[o:].data_ptr(), iters[o:].data_ptr(), llr.data_ptr(), _lib.LLR_FAILED,
                                                 fail[4:].data_ptr(), fail.data_ptr(), iter_total.data_ptr(), stream), "bp")
                e1.record()
                _lib.check(L.qldpc_osd_decode_dev(code.handle, fail[4:].data_ptr(), fail.data_ptr(), b, synd[o:].data_ptr(), llr.data_ptr(),
                                                  f64, corr[o:].data_ptr(), corr[o:].data_ptr(), valid[o:].data_ptr(), stream), "osd")
                e2.record()
                # performOSD_enhanced(order=7): sweep on the OSD-0 solutions that miss the syndrome (device-side list; BP hard
                # decision = llr < 0 because the OSD-0 call wrote its solution over it)
                _lib.check(L.qldpc_osdw_decode_dev(code.handle, fail[4:].data_ptr(), fail.data_ptr(), b, synd[o:].data_ptr(), llr.data_ptr(),
                                                   f64, None, corr[o:].data_ptr(), valid[o:].data_ptr(), OSD_ORDER, 0, stream), "osdw")
                e3.record()
                marks.append((e0, e1, e2, e3))
                launches += 4            # BP, OSD-0, compaction of the invalid shots, sweep (memsets not counted)
            e4, e5 = ev(), ev()
            e4.record()
            _lib.check(L.qldpc_check_dev(code.handle, B, err.data_ptr(), corr.data_ptr(), synd.data_ptr(), conv.data_ptr(),
                                         iters.data_ptr(), distance, None, None, counters.data_ptr(), stream), "check")
            e5.record()
            launches += 1
            return marks, (e4, e5)

        for _ in range(args.warmup):
            step()
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
            all_marks.append(step())
        cnt = counters.clone()
        if world > 1:
            dist.all_reduce(cnt, op=dist.ReduceOp.SUM)      # the one data-path collective: LER counters (inside the timed region)
        t_end.record()
        barrier()
        clocks = sampler.stop()
        total_ms = t_start.elapsed_time(t_end)
        for marks, (e4, e5) in all_marks:
            for e0, e1, e2, e3 in marks:
                kernel_ms["bp"] += e0.elapsed_time(e1)
                kernel_ms["osd0"] += e1.elapsed_time(e2)
                kernel_ms["osdw_sweep_stage"] += e2.elapsed_time(e3)
            kernel_ms["check"] += e4.elapsed_time(e5)
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        itot = iter_total.clone()
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.all_reduce(itot, op=dist.ReduceOp.SUM)
        total_ms = float(t.item())
        value = B * args.steps * world / (total_ms * 1e-3)
        cd = dict(zip(_lib.COUNTER_NAMES, (int(x) for x in cnt.cpu().numpy())))

        # ---------------- e2e: host buffers through the C ABI ----------------
        def e2e_step():
            _lib.check(L.qldpc_bposd_decode_host(code.handle, ctypes.byref(cfg), prior_p, Be, synd_h.data_ptr(), OSD_ORDER,
                                                 corr_h.data_ptr(), conv_h.data_ptr(), iters_h.data_ptr()), "bposd_host")

        def e2e_packed_step():
            _lib.check(L.qldpc_bposd_decode_host_packed(code.handle, ctypes.byref(cfg), prior_p, Be, synd_hp.data_ptr(), OSD_ORDER,
                                                        corr_hp.data_ptr(), conv_h.data_ptr(), iters_h.data_ptr()), "bposd_host_packed")

        def wall(fn, reps):
            fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(reps):
                fn()
            torch.cuda.synchronize()
            tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            return float(tt.item())

        e2e_steps = max(1, min(args.steps, 3))
        corr_u8_dev = torch.empty((Be, n), dtype=torch.uint8, device=dev)
        _lib.check(L.qldpc_unpack_bits_dev(corr.data_ptr(), corr_u8_dev.data_ptr(), Be, n, stream), "unpack")
        torch.cuda.synchronize()
        corr_u8_ref = corr_u8_dev.cpu()
        del corr_u8_dev

        def e2e_leg(mode):
            """the uint8 host call with the rows packed by the library's own choice (-1), on the device (0) or by host threads (1):
            shots/s, bytes per step as counted by the library from the copies it issued, equality with the device path"""
            code.set_host_pack(mode)
            corr_h.zero_()
            s0 = code.host_transfer_stats()
            v = Be * e2e_steps * world / wall(e2e_step, e2e_steps)
            s1 = code.host_transfer_stats()
            calls = e2e_steps + 1
            names = {0: "device kernels (byte rows cross PCIe)", 1: "host threads (bit-packed rows cross PCIe)",
                     2: "chunk by chunk, whichever side is free (qldpc_b200.h: qldpc_host_transfer_stats)"}
            return dict(value=v, unit="shots/s", h2d_bytes_per_step=(s1["h2d_bytes"] - s0["h2d_bytes"]) // calls,
                        d2h_bytes_per_step=(s1["d2h_bytes"] - s0["d2h_bytes"]) // calls,
                        rows_packed_by=names[s1["host_pack"]], chunks_packed_by_host=s1["chunks_host"] - s0["chunks_host"],
                        chunks_packed_by_device=s1["chunks_device"] - s0["chunks_device"],
                        host_pack_rate_shots_per_s=s1["host_pack_rate"], matches_device_path=bool(torch.equal(corr_u8_ref, corr_h)))
        e2e_auto = e2e_leg(-1)
        e2e_other = {"device": e2e_leg(0), "host": e2e_leg(1), "chunk_by_chunk": e2e_leg(2)}
        code.set_host_pack(-1)
        e2e_value, e2e_matches = e2e_auto["value"], e2e_auto["matches_device_path"]
        del corr_u8_ref
        e2e_packed_value = Be * e2e_steps * world / wall(e2e_packed_step, e2e_steps)
        e2e_packed_matches = bool(torch.equal(corr_hp, corr[:Be].cpu()))
        # the reference's actual Monte-Carlo use: everything on the device, counters back (no per-shot host traffic)
        mc_counters = np.zeros(_lib.NUM_COUNTERS, np.uint64)

        def mc_step():
            _lib.check(L.qldpc_mc_sweep(code.handle, ctypes.byref(cfg), prior_p, P_ERR, SEED, rank * B, B, 1, OSD_ORDER, distance,
                                        mc_counters.ctypes.data_as(ctypes.c_void_p)), "mc_sweep")
        mc_value = B * e2e_steps * world / wall(mc_step, e2e_steps)
        mc = dict(zip(_lib.COUNTER_NAMES, (int(x) // (e2e_steps + 1) for x in mc_counters)))

        out = {"value": value, "unit": "shots/s", "ms_per_step": total_ms / args.steps, "dtype": "f%d" % precision,
               "kernel_ms_per_step": {k: v / args.steps for k, v in kernel_ms.items()},
               "results": {"ler": cd["logical"] / max(1, cd["shots"]), "bp_failure_rate": cd["bp_failed"] / max(1, cd["shots"]),
                           "invalid": cd["invalid"], "mean_exit_iteration": cd["iter_sum"] / max(1, cd["shots"]), "shots": cd["shots"]},
               "e2e": {"value": e2e_value, "unit": "shots/s", "h2d_bytes_per_step": e2e_auto["h2d_bytes_per_step"],
                       "d2h_bytes_per_step": e2e_auto["d2h_bytes_per_step"],
                       "shots_per_step": Be, "steps": e2e_steps, "api": "qldpc_bposd_decode_host (uint8 in pinned host memory), osd_order=7",
                       "matches_device_path": e2e_matches, "rows_packed_by": e2e_auto["rows_packed_by"],
                       "chunks_packed_by_host": e2e_auto["chunks_packed_by_host"], "chunks_packed_by_device": e2e_auto["chunks_packed_by_device"],
                       "host_pack_rate_shots_per_s": e2e_auto["host_pack_rate_shots_per_s"],
                       "host_bytes_per_step": {"in": Be * m, "out": Be * (n + 1 + 4)},
                       "forced_modes": e2e_other,
                       "packed_host_rows": {"value": e2e_packed_value, "unit": "shots/s", "h2d_bytes_per_step": Be * 4 * WM,
                                            "d2h_bytes_per_step": Be * (4 * WN + 1 + 4), "api": "qldpc_bposd_decode_host_packed",
                                            "matches_device_path": e2e_packed_matches},
                       "mc_sweep": {"value": mc_value, "unit": "shots/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * _lib.NUM_COUNTERS,
                                    "api": "qldpc_mc_sweep (device Philox sampler -> BP -> OSD -> checks -> counters)",
                                    "ler": mc["logical"] / max(1, mc["shots"]),
                                    "counters_match_device_path": (mc["logical"] * args.steps == cd["logical"]) if world == 1 else None}},
               "gpu_launches": launches, "clocks": clocks, "geometry": geom}
        if rank == 0:
            E = code.E
            A = 15 * E + 2 * n + m                            # lane-ops per shot-iteration
            iters_exec = int(itot.item()) / world             # per rank, over the timed steps
            bp_s = kernel_ms["bp"] * 1e-3
            achieved = iters_exec * A / bp_s
            llr_bytes = (8 if f64 else 4) * n
            hbm_bytes_per_shot = 4 * WM + 4 * WN + 1 + 4 + llr_bytes * cd["bp_failed"] / max(1, cd["shots"])
            kname = ("bp_warp_kernel_f64<CPL=%d,VPL=%d,RW=6>" if f64 else "bp_warp_kernel<CPL=%d,VPL=%d,RW=6>") % (WM, WN)
            roof = {"bound": "alu", "kernel": kname if geom["kernel"] == "warp_per_shot" else geom["kernel"],
                    "achieved": achieved / 1e12, "peak": peak_laneops / 1e12, "unit": "Tlane-op/s", "frac": achieved / peak_laneops,
                    "peak_def": f"{n_sm} SMs x 128 lanes x {sm_max:.0f} MHz (max clock; median clock under this arm {clocks.get('sm_mhz') or 0:.0f} MHz)",
                    "algorithmic_unit": f"A = 15E+2n+m = {A} lane-ops per shot-iteration; {iters_exec / (B * args.steps):.2f} iterations/shot executed",
                    "kernel_ms_per_step": kernel_ms["bp"] / args.steps, "shot_iterations_per_s": iters_exec / bp_s,
                    "hbm": {"algorithmic_bytes_per_shot": hbm_bytes_per_shot, "achieved_gbs": hbm_bytes_per_shot * B * args.steps / bp_s / 1e9,
                            "peak_gbs": peaks.get("hbm_gbs"), "note": "on-chip path: HBM traffic is negligible by design"},
                    "traffic": None}
            cap_file = "r2_bp_f64_ncu.json" if f64 else "r1k_bp_ncu.json"
            try:   # DRAM bytes per launch and pipe utilisation from the committed ncu capture of the same kernel
                cap = json.load(open(os.path.join(ROOT, "profiles", cap_file)))
                roof["traffic"] = cap["dram_bytes_per_shot"] * min(B, CH)
                roof["ncu_capture"] = cap
            except Exception:
                pass
            out["roofline"] = roof
        return out

    arm32 = run_arm(32)
    arm64 = run_arm(64)

    # ---------------- bit-exactness of the float64 arm / divergence of the float32 arm against the CPU port ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        synd_cpu = synd_h.numpy()
        v, cores, sample, dt, r = cpu_port_throughput(H, synd_cpu, prior)
        cfg64, cfg32 = Code.config(precision=64, **BP), Code.config(precision=32, **BP)
        same = {}
        for name, cf in (("f64", cfg64), ("f32", cfg32)):
            _lib.check(L.qldpc_bposd_decode_host(code.handle, ctypes.byref(cf), prior_p, sample, synd_h.data_ptr(), OSD_ORDER,
                                                 corr_h.data_ptr(), conv_h.data_ptr(), iters_h.data_ptr()), "bposd_host")
            same[name] = float((r["corr"] == corr_h[:sample].numpy()).all(1).mean())
            if name == "f64":
                same["f64_flags_and_iterations_identical"] = bool((r["converged"] == conv_h[:sample].numpy().astype(bool)).all() and
                                                                  (r["iters"] == iters_h[:sample].numpy()).all())
                corr64 = corr_h[:4096].numpy().copy()
        cpu = {"value": v, "unit": "shots/s", "cores": cores, "kind": "port",
               "sample": f"first {sample} shots of the same workload ({dt:.1f} s; float64 C port of min-sum BP + OSD, OpenMP)",
               "fraction_of_shots_with_identical_correction_vs_gpu_f32": same["f32"],
               "fraction_of_shots_with_identical_correction_vs_gpu_f64": same["f64"],
               "gpu_f64_flags_and_iterations_identical": same["f64_flags_and_iterations_identical"]}
        rp = reference_python_throughput(H, synd_cpu, prior, check=corr64)
        if rp:
            cpu["reference_python"] = rp

    extras = {}
    if not args.no_extras:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import bench_extras as X
        # ---------------- sustained: BASELINE config 3, 10^9 shots over the ranks, one all-reduce of the counters ----------------
        H3, L3, d3 = load_code_arrays("[[288, 12, 18]]")
        c3 = Code(H3, L3, graph.reference_schedule(H3, "min_sum"), d3)
        cfg3 = Code.config(precision=32, variant="min_sum", max_iter=50, alpha=0.8, damping=0.7, clip=25.0)
        p3 = 0.06
        pr3 = np.full(H3.shape[1], np.log((1 - p3) / p3))
        S3 = args.sustained_shots
        lo, hi = S3 * rank // world, S3 * (rank + 1) // world
        cnt3 = np.zeros(_lib.NUM_COUNTERS, np.uint64)
        _lib.check(L.qldpc_mc_sweep(c3.handle, ctypes.byref(cfg3), pr3.ctypes.data_as(ctypes.c_void_p), p3, SEED, lo, min(hi - lo, 1 << 22), 1, -1,
                                    d3, cnt3.ctypes.data_as(ctypes.c_void_p)), "mc_sweep warm-up")
        cnt3[:] = 0
        barrier()
        sampler = ClockSampler(local_rank)
        sampler.start()
        time.sleep(0.3)
        t0 = time.perf_counter()
        _lib.check(L.qldpc_mc_sweep(c3.handle, ctypes.byref(cfg3), pr3.ctypes.data_as(ctypes.c_void_p), p3, SEED, lo, hi - lo, 1, -1, d3,
                                    cnt3.ctypes.data_as(ctypes.c_void_p)), "mc_sweep")
        ct = torch.from_numpy(cnt3.astype(np.int64)).to(dev)
        if world > 1:
            dist.all_reduce(ct, op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        clk3 = sampler.stop()
        c3d = dict(zip(_lib.COUNTER_NAMES, (int(x) for x in ct.cpu().numpy())))
        A3 = 15 * c3.E + 2 * c3.n + c3.m
        extras["sustained"] = {
            "config": "3: [[288,12,18]] BP-only (min-sum 0.8/0.7/25, <= 50 iterations, float32), p=0.06, %d shots sharded over %d GPU(s), device sampler, one all-reduce of the counters" % (S3, world),
            "seconds": float(tt.item()), "shots_per_s": S3 / float(tt.item()), "shots": c3d["shots"],
            "bp_failures": c3d["bp_failed"], "logical_on_bp_converged": c3d["logical_and_bp_converged"],
            "mean_exit_iteration": c3d["iter_sum"] / max(1, c3d["shots"]),
            "lane_op_frac_whole_pipeline": (c3d["iter_sum"] + c3d["shots"]) * A3 / float(tt.item()) / (peak_laneops * world),
            "clocks": clk3}
        del c3
        # ---------------- the other BASELINE configs, single GPU figures (rank 0) ----------------
        if rank == 0:
            extras["configs"] = X.baseline_configs(quick=True)
    if rank == 0:
        line = {"metric": METRIC, "value": arm32["value"], "unit": "shots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": arm32["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": dict(base_config(B),
                               l2="per-step inputs+outputs (%.0f MB) exceed the 126 MB L2" % ((4 * WM + 8 * WN + 5) * B / 1e6),
                               geometry=arm32.pop("geometry"), parallelism=f"shots sharded over {world} GPU(s), one all-reduce of counters inside the timed region"),
                "kernel_ms_per_step": arm32["kernel_ms_per_step"], "results": arm32["results"], "roofline": arm32["roofline"],
                "e2e": arm32["e2e"], "gpu_launches": arm32["gpu_launches"], "clocks": arm32["clocks"],
                "bit_exact_f64": arm64}
        line.update(extras)
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
