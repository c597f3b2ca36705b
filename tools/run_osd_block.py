"""One launch of the block-per-shot OSD-0 kernel on the space-time matrix (864 x 2592) -- profiling target.
    python tools/run_osd_block.py [B]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from qldpc_b200 import Code, _lib, graph
from qldpc_b200.spaceTime import spaceTimeMatrix
B = int(sys.argv[1]) if len(sys.argv) > 1 else 592
d = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qldpc_b200", "data", "codes", "[[144, 12, 12]].npz"))
Hst = spaceTimeMatrix(d["Hx"], 12)
code = Code(Hst, None, (graph.SEQ, graph.SEQ))
rng = np.random.default_rng(1)
L = _lib.lib(); dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
llr = torch.from_numpy(rng.normal(3.0, 2.5, (B, code.n)).astype(np.float32)).to(dev)
hard = torch.from_numpy((rng.random((B, code.words_n)) < 0).astype(np.int32)).to(dev)
err = (rng.random((B, code.n)) < 0.01).astype(np.uint8)
synd_u8 = torch.from_numpy(code.syndromes(err).astype(np.uint8)).to(dev)
synd = torch.zeros((B, code.words_m), dtype=torch.int32, device=dev)
_lib.check(L.qldpc_pack_bits_dev(synd_u8.data_ptr(), synd.data_ptr(), B, code.m, st))
out = torch.empty((B, code.words_n), dtype=torch.int32, device=dev)
valid = torch.empty(B, dtype=torch.uint8, device=dev)
def run():
    _lib.check(L.qldpc_osd_decode_dev(code.handle, None, None, B, synd.data_ptr(), llr.data_ptr(), 0, hard.data_ptr(), out.data_ptr(), valid.data_ptr(), st))
run(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record(); torch.cuda.synchronize()
print(f"B={B}: {e0.elapsed_time(e1):.1f} ms  -> {B/e0.elapsed_time(e1)*1e3:.0f} OSD shots/s; valid {int(valid.sum())}/{B}")
