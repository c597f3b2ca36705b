"""One HBM-staged BP launch on the space-time matrix (BASELINE config 4) -- profiling target.
    python tools/run_staged.py [B] [p] [reps]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qldpc_b200 import Code, _lib, graph
from qldpc_b200.spaceTime import spaceTimeMatrix
B = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.005
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qldpc_b200", "data", "codes", "[[144, 12, 12]].npz"))
H = d["Hx"]; m, n = H.shape
Hst = spaceTimeMatrix(H, 12)
code = Code(Hst, None, (graph.SEQ, graph.SEQ))
rng = np.random.default_rng(4)
err = (rng.random((B, n)) < p).astype(np.int64); s = (err @ H.T) % 2; hist = []
for _ in range(12):
    s = (s + (rng.random((B, m)) < p)) % 2; hist.append(s)
synd = np.concatenate([hist[-1]] + [(hist[i] + hist[i - 1]) % 2 for i in range(1, 12)], axis=1).astype(np.uint8)
L = _lib.lib(); dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
u8 = torch.from_numpy(synd).to(dev)
sp = torch.zeros((B, code.words_m), dtype=torch.int32, device=dev)
_lib.check(L.qldpc_pack_bits_dev(u8.data_ptr(), sp.data_ptr(), B, code.m, st))
hard = torch.empty((B, code.words_n), dtype=torch.int32, device=dev); conv = torch.empty(B, dtype=torch.uint8, device=dev)
iters = torch.empty(B, dtype=torch.int32, device=dev); itot = torch.zeros(1, dtype=torch.int64, device=dev)
cfg = Code.config("min_sum", 50, 0.8, 0.7, 25.0, 32)
prior = np.full(code.n, np.log((1 - p) / p))
def run():
    _lib.check(L.qldpc_bp_decode_dev(code.handle, ctypes.byref(cfg), prior.ctypes.data_as(ctypes.c_void_p), B, sp.data_ptr(), hard.data_ptr(),
                                     conv.data_ptr(), iters.data_ptr(), None, 0, None, None, itot.data_ptr(), st))
run(); torch.cuda.synchronize(); itot.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
it = itot.item() / reps
print(f"B={B} p={p}: {ms:.1f} ms, {it/B:.2f} it/shot, {it/ms*1e3:.3e} shot-iter/s, {it*12*code.E/ms*1e3/1e9:.0f} GB/s algorithmic ({it*12*code.E/ms*1e3/1e9/6546.9:.3f} of measured HBM peak)")
