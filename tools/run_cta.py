"""One launch of the CTA-per-shot BP kernel on the space-time matrix (864 x 2592) -- profiling target.
    python tools/run_cta.py [B] [p]"""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qldpc_b200 import Code, _lib, graph
from qldpc_b200.spaceTime import spaceTimeMatrix

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.003
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qldpc_b200", "data", "codes", "[[144, 12, 12]].npz"))
Hst = spaceTimeMatrix(d["Hx"], 12)
code = Code(Hst, None, (graph.SEQ, graph.SEQ))
L = _lib.lib(); dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(1)
err = (rng.random((B, code.n)) < p).astype(np.uint8)
synd_u8 = torch.from_numpy(code.syndromes(err).astype(np.uint8)).to(dev)
synd = torch.zeros((B, code.words_m), dtype=torch.int32, device=dev)
_lib.check(L.qldpc_pack_bits_dev(synd_u8.data_ptr(), synd.data_ptr(), B, code.m, st))
hard = torch.empty((B, code.words_n), dtype=torch.int32, device=dev)
conv = torch.empty(B, dtype=torch.uint8, device=dev)
iters = torch.empty(B, dtype=torch.int32, device=dev)
itot = torch.zeros(1, dtype=torch.int64, device=dev)
cfg = Code.config("min_sum", 50, 0.8, 0.7, 25.0, 32)
prior = np.full(code.n, np.log((1 - p) / p))
print(code.geometry(cfg))
def run():
    _lib.check(L.qldpc_bp_decode_dev(code.handle, ctypes.byref(cfg), prior.ctypes.data_as(ctypes.c_void_p), B, synd.data_ptr(), hard.data_ptr(),
                                     conv.data_ptr(), iters.data_ptr(), None, 0, None, None, itot.data_ptr(), st))
run(); torch.cuda.synchronize(); itot.zero_()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"B={B} p={p}: {ms:.2f} ms, {itot.item() / ms * 1e3 / 1e6:.1f} M shot-iterations/s, converged {conv.float().mean().item():.3f}")
