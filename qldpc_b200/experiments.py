"""Monte-Carlo drivers that emit the reference's result dictionaries, so that its
plotting code / loadResults.py can consume them unchanged:

    paper_results()      -> keys of paperResults.py:110-114 / paperResults_GPU.py:156-160 / degeneracyCount.ipynb
    rework_main()        -> keys of rework/main.py:119-129 (and rework/Alvarado.py:196-206)
    rework_main_different_orders() -> results[code][label][p] of rework/main_different_orders.py:45-50,125-135
    bp_per_iteration()   -> keys of BP_per_Iteration.py:85-90
    save_results()       -> np.savez(path, results=dict) as every reference script does

Counting runs entirely on the device (Code.mc_sweep: Philox sampling -> BP -> OSD -> checks -> counters); drivers that
must return per-shot lists (residual weights, posterior LLRs) use the batched host-array calls.  With
torch.distributed initialised, shot ranges are sharded over the ranks and the counters all-reduced once.

Random streams: every (code, p) point of a driver gets its own range of global shot ids of the Philox stream `seed` (the
k-th point uses ids [k * trials, (k + 1) * trials)), so the points of a curve are independent, as in the reference, which
draws them one after the other from one NumPy stream (paperResults.py:61-63).
"""
import os

import numpy as np

from . import _lib
from .code import Code, DATA_CODES, load_code
from . import graph as _graph

CODES = ["[[72, 12, 6]]", "[[90, 8, 10]]", "[[108, 8, 10]]", "[[144, 12, 12]]", "[[288, 12, 18]]"]


# ---- sharding over ranks (the only multi-GPU logic there is) -------------------------------------
def shard_range(nshots, rank, world):
    """Contiguous block of global shot ids [first, first + count) decoded by `rank`."""
    base, rem = divmod(int(nshots), int(world))
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def _dist():
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist
    except Exception:
        pass
    return None


def allreduce_counters(counters):
    """Sum a dict of integer counters over all ranks with ONE all-reduce (NCCL on GPUs, gloo in CPU tests)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return dict(counters)
    import torch
    keys = list(counters)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([int(counters[k]) for k in keys], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return {k: int(v) for k, v in zip(keys, t.cpu().tolist())}


def mc_point(code, p, nshots, seed=0, draws=1, first_shot=0, **decoder):
    """One (code, p) Monte-Carlo point on the global shot ids [first_shot, first_shot + nshots): this rank's shard of
    that range, counters reduced over ranks."""
    dist = _dist()
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist else (0, 1)
    first, count = shard_range(nshots, rank, world)
    c = (code.mc_sweep(p, count, seed=seed, first_shot=first_shot + first, draws=draws, **decoder) if count
         else dict.fromkeys(_lib.COUNTER_NAMES, 0))
    return allreduce_counters(c)


# ---- paperResults.py / paperResults_GPU.py / degeneracyCount.ipynb ---------------------------------
def paper_results(codes=CODES, physicalErrorRates=(0.05, 0.04, 0.03, 0.02, 0.01, 0.009, 0.008, 0.007), trials=1000,
                  variant="sum_product", maxIter=200, osd_order=0, draws=2, seed=0, precision=64, bp_only=False,
                  directory=None, **bp_kwargs):
    """paperResults.py:33-114.  draws=2 is its error model (XOR of two Bernoulli(p) draws, :61-63); the prior is
    ln((1-p)/p) of the NOMINAL p as in :49.  bp_only=True gives the BP-only accounting of degeneracyCount.ipynb
    cell 5 (a BP failure counts as a logical error and the logical check of the failed detection is added on top)."""
    results = {}
    point = 0
    for name in codes:
        code = load_code(name, "x", directory)
        out = {k: [] for k in ("ler", "BPs_fault", "BPs_miscorrected", "incorrectable", "degeneracies")}
        for p in physicalErrorRates:
            prior = np.log((1 - p) / p)
            c = mc_point(code, p, trials, seed=seed, draws=draws, first_shot=point * trials, prior=prior, variant=variant,
                         max_iter=maxIter, osd_order=(-1 if bp_only else osd_order), precision=precision, **bp_kwargs)
            point += 1
            logical = c["logical"] + (c["bp_failed"] if bp_only else 0)
            out["ler"].append(logical / trials)
            out["BPs_fault"].append(c["bp_failed"] if bp_only else 0)     # paperResults.py leaves this counter at 0 (:74-75)
            out["BPs_miscorrected"].append(c["miscorrected"])
            out["incorrectable"].append(c["incorrectable"])
            out["degeneracies"].append(c["degenerate"])
        results[name] = out
    return results


# ---- rework/main.py / rework/Alvarado.py ------------------------------------------------------------
def rework_main(experiment, trials=10000, BP_maxIter=100, OSD_order=7, variant="sum_product", seed=0, precision=64,
                directory=None, chunk=1 << 20, alpha=1.0, damping=1.0, clip=20.0, first_point=0):
    """rework/main.py:51-129.  `experiment` is its list of dicts {"code", "name", "physicalErrorRates", "distance"}.
    Per-shot lists (residual weights by category) are gathered from the batched calls."""
    results = {}
    point = first_point
    for exp in experiment:
        code = load_code(exp["code"], "x", directory)
        results[exp["name"]] = {}
        for p in exp["physicalErrorRates"]:
            prior = np.log((1 - p) / p)
            acc = dict(logical=0, osd=0, degen=0, both=0, iters=0)
            w_bp, w_osd, w_bp_err, w_osd_err = [], [], [], []
            for o in range(0, trials, chunk):
                b = min(chunk, trials - o)
                err, synd = code.sample(p, b, seed=seed, first_shot=point * trials + o)
                corr, conv, iters = code.bposd_decode_batch(synd, prior, variant, BP_maxIter, alpha, damping, clip,
                                                            precision=precision, osd_order=OSD_order)
                chk = code.check_batch(err, corr, synd, conv, iters)
                lg, va, wt = chk["logical"], chk["valid"], chk["weight"]
                differs = wt > 0
                acc["logical"] += int(lg.sum()); acc["osd"] += int((~conv).sum()); acc["iters"] += int(iters.sum())
                acc["both"] += int((lg & ~conv).sum()); acc["degen"] += int((~lg & differs & va).sum())
                w_osd_err += wt[lg & ~conv].tolist(); w_bp_err += wt[lg & conv].tolist()
                w_osd += wt[~lg & differs & ~conv].tolist(); w_bp += wt[~lg & differs & conv].tolist()
            results[exp["name"]][p] = {
                "logical": acc["logical"] / trials, "osd": acc["osd"] / trials, "degeneracies": acc["degen"] / trials,
                "average_iterations": acc["iters"] / trials, "OSD_invocation_AND_logicalError": acc["both"] / trials,
                "weights_found_BP": w_bp, "weights_found_OSD": w_osd,
                "weights_found_BP_error": w_bp_err, "weights_found_OSD_error": w_osd_err}
            point += 1
    return results


DIFFERENT_ORDERS = ({"bp_iter": 50, "osd_order": 0, "label": "BP50_OSD0"}, {"bp_iter": 100, "osd_order": 0, "label": "BP100_OSD0"},
                    {"bp_iter": 50, "osd_order": 7, "label": "BP50_OSD7"}, {"bp_iter": 100, "osd_order": 7, "label": "BP100_OSD7"})


def rework_main_different_orders(experiment, configurations=DIFFERENT_ORDERS, trials=10000, **kwargs):
    """rework/main_different_orders.py:45-135: the rework/main.py loop for BP{50,100} x OSD{0,7};
    results[code name][configuration label][p] = the rework/main.py entry (saved as simulation_results_complex.npz)."""
    results = {exp["name"]: {} for exp in experiment}
    point = 0
    for exp in experiment:
        for cfg in configurations:
            r = rework_main([exp], trials=trials, BP_maxIter=cfg["bp_iter"], OSD_order=cfg["osd_order"], first_point=point, **kwargs)
            for entry in r[exp["name"]].values():
                entry.pop("average_iterations")            # (not stored by main_different_orders.py:125-134)
            results[exp["name"]][cfg["label"]] = r[exp["name"]]
            point += len(exp["physicalErrorRates"])
    return results


# ---- BP_per_Iteration.py ---------------------------------------------------------------------------
def bp_per_iteration(codes=CODES, errorRate=0.01, iterations=(10, 20, 30, 40, 50, 60, 70, 80, 90), trials=10000,
                     variant="sum_product", seed=0, precision=64, directory=None, alpha=1.0, damping=1.0, clip=20.0,
                     keep_llrs=True):
    """BP_per_Iteration.py:27-90: LER / degeneracy / OSD-invocation rate per BP iteration budget, plus every posterior
    LLR (`llrs_per_iter`) and the LLRs of the BP-failed shots (`llrs_per_iter_after_OSD`) as flat float64 arrays."""
    results = {}
    for name in codes:
        code = load_code(name, "x", directory)
        prior = np.log((1 - errorRate) / errorRate)
        out = dict(logicalErrors=[], degeneracies=[], OSD_invocations=[], iterations=list(iterations), llrs_per_iter=[],
                   llrs_per_iter_after_OSD=[])
        for k, max_iter in enumerate(iterations):
            err, synd = code.sample(errorRate, trials, seed=seed, first_shot=k * trials)
            hard, conv, llr, iters = code.bp_decode_batch(synd, prior, variant, max_iter, alpha, damping, clip,
                                                          precision=precision, want_llr=True)
            corr = hard.astype(np.uint8)
            f = np.nonzero(~conv)[0]
            if f.size:
                corr[f] = code.osd_decode_batch(synd[f], llr[f], hard[f]).astype(np.uint8)
            chk = code.check_batch(err, corr, synd, conv, iters)
            out["logicalErrors"].append(float(chk["logical"].mean()))
            out["degeneracies"].append(float(chk["degenerate"].mean()))
            out["OSD_invocations"].append(float((~conv).mean()))
            out["llrs_per_iter"].append(llr.reshape(-1) if keep_llrs else np.zeros(0))
            out["llrs_per_iter_after_OSD"].append(llr[f].reshape(-1) if keep_llrs else np.zeros(0))
        results[name] = out
    return results


def save_results(path, results):
    """np.savez(path, results=dict): what every reference script writes and loadResults.py reads back with .item()."""
    np.savez(path, results=np.array(results, dtype=object))        # (object arrays are pickled by default: one key, `results`)
