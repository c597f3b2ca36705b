"""Times qldpc_bp_decode_dev on device-resident syndromes for a sweep of launch geometries.
    python tools/tune_bp.py [--code "[[144, 12, 12]]"] [--p 0.05] [--shots 4000000]"""
import argparse, ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qldpc_b200 import Code, _lib, graph

ap = argparse.ArgumentParser()
ap.add_argument("--code", default="[[144, 12, 12]]")
ap.add_argument("--p", type=float, default=0.05)
ap.add_argument("--shots", type=int, default=4_000_000)
ap.add_argument("--max-iter", type=int, default=100)
ap.add_argument("--configs", default="0:0:0,3:0:0,0:4:1,0:8:1,0:8:2,2:0:0")
args = ap.parse_args()
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qldpc_b200", "data", "codes", args.code + ".npz"))
H = d["Hx"]
m, n = H.shape
code = Code(H, d["Lx"], (graph.SEQ, graph.SEQ), int(d["distance"]))
L = _lib.lib()
dev = torch.device("cuda", 0)
B = args.shots
WM, WN = code.words_m, code.words_n
err = torch.empty((B, WN), dtype=torch.int32, device=dev)
synd = torch.empty((B, WM), dtype=torch.int32, device=dev)
hard = torch.empty((B, WN), dtype=torch.int32, device=dev)
conv = torch.empty(B, dtype=torch.uint8, device=dev)
iters = torch.empty(B, dtype=torch.int32, device=dev)
itot = torch.zeros(1, dtype=torch.int64, device=dev)
st = torch.cuda.current_stream().cuda_stream
_lib.check(L.qldpc_sample_dev(code.handle, args.p, 0, 0, 1, B, err.data_ptr(), synd.data_ptr(), st))
prior = np.full(n, np.log((1 - args.p) / args.p))
ref = None
try:
    print('conflict model (wavefronts per warp-iteration of the variable pass) T=4:', code.tiled_conflict_model(4), 'T=8:', code.tiled_conflict_model(8))
except Exception as ex:
    print('no conflict model:', ex)
for spec in args.configs.split(","):
    staged, T, rmin = (int(x) for x in spec.split(":"))
    cfg = Code.config("min_sum", args.max_iter, 0.8, 0.7, 25.0, 32, staged, T, rmin)
    geom = code.geometry(cfg)
    def run():
        _lib.check(L.qldpc_bp_decode_dev(code.handle, ctypes.byref(cfg), prior.ctypes.data_as(ctypes.c_void_p), B, synd.data_ptr(),
                                         hard.data_ptr(), conv.data_ptr(), iters.data_ptr(), None, 0, None, None, itot.data_ptr(), st))
    run(); torch.cuda.synchronize(); itot.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    sig = (int(hard.sum().item()), int(conv.sum().item()), int(iters.sum().item()))
    if ref is None: ref = sig
    print(f"staged={staged} T={T} refill_min={rmin} kernel={geom['kernel']} lanes={geom['lanes_per_shot']} shots/cta={geom['shots_per_cta']} "
          f"smem={geom['smem_bytes']}: {ms:8.2f} ms  {B/ms*1e3/1e6:7.2f} Mshots/s  {itot.item()/2/ms*1e3/1e9:6.3f} Gshot-iter/s  same={sig==ref}")
