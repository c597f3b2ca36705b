"""Concurrent host<->device bandwidth of all ranks of one box (torchrun, one rank per GPU): every rank copies its own pinned
buffers D2H, H2D and both at once while the others do the same; prints per-rank and aggregate GB/s.  QLDPC_PIN_LOCAL=1 binds the
process to the CPUs nvidia-smi lists as local to its GPU before the pinned buffers are allocated (first touch)."""
import os, subprocess, time, json
import torch
import torch.distributed as dist

rank = int(os.environ.get("RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1)); local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
aff = None
if os.environ.get("QLDPC_PIN_LOCAL") == "1":
    try:
        t = subprocess.run(["nvidia-smi", "topo", "-C", "-i", str(local)], capture_output=True, text=True).stdout
        # fall back to the matrix
        m = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
        import re
        for line in m.splitlines():
            if re.match(r"^(\x1b\[\d*m)?GPU%d\s" % local, line):
                cols = re.sub(r"\x1b\[\d*m", "", line).split("\t")
                cand = [c.strip() for c in cols if re.fullmatch(r"[\d,\-]+", c.strip() or "x") and ("-" in c or "," in c)]
                if cand:
                    cpus = set()
                    for part in cand[0].split(","):
                        a, _, b = part.partition("-")
                        cpus.update(range(int(a), int(b or a) + 1))
                    os.sched_setaffinity(0, cpus)
                    aff = cand[0]
    except Exception as e:
        aff = "failed: %r" % e
N_OUT, N_IN = 745_000_000, 360_000_000
d = torch.empty(N_OUT, dtype=torch.uint8, device="cuda"); h = torch.empty(N_OUT, dtype=torch.uint8).pin_memory(); h.zero_()
d2 = torch.empty(N_IN, dtype=torch.uint8, device="cuda"); h2 = torch.empty(N_IN, dtype=torch.uint8).pin_memory(); h2.zero_()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

def timed(fn, reps=4):
    fn(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / reps
    x = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1: dist.all_reduce(x, op=dist.ReduceOp.MAX)
    return dt, float(x.item())

def f_d2h(): h.copy_(d, non_blocking=True)
def f_h2d(): d2.copy_(h2, non_blocking=True)
def f_both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
res = {}
for name, fn, nbytes in (("d2h", f_d2h, N_OUT), ("h2d", f_h2d, N_IN), ("both", f_both, N_OUT + N_IN)):
    dt, dtmax = timed(fn)
    res[name] = dict(rank_gbs=nbytes / dt / 1e9, aggregate_gbs=world * nbytes / dtmax / 1e9)
allres = [None] * world
if world > 1:
    dist.all_gather_object(allres, dict(rank=rank, aff=aff, **{k: round(v["rank_gbs"], 1) for k, v in res.items()}))
else:
    allres = [dict(rank=0, aff=aff, **{k: round(v["rank_gbs"], 1) for k, v in res.items()})]
if rank == 0:
    print(json.dumps(dict(world=world, pin_local=os.environ.get("QLDPC_PIN_LOCAL"), aggregate_gbs={k: round(v["aggregate_gbs"], 1) for k, v in res.items()}, ranks=allres)))
if world > 1:
    dist.destroy_process_group()
