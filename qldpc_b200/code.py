"""`Code`: a parity-check matrix resident on the GPU (Tanner-graph tables, packed columns, packed
logical operators, workspaces) plus the batched decode calls on it.

This is the host-side mirror of the C ABI (include/qldpc_b200.h); the reference-named functions in
qldpc_b200.decoding.* / qldpc_b200.rework.decoding are thin B = 1 wrappers over these methods.
All heavy lifting happens in libqldpc_b200.so; nothing here computes a decode on the CPU.
"""
import ctypes
import hashlib
import os

import numpy as np

from . import _lib
from . import graph as _graph

VARIANTS = {"min_sum": _lib.MIN_SUM, "sum_product": _lib.SUM_PRODUCT, "sum_product_sym": _lib.SUM_PRODUCT_SYM}
DATA_CODES = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "codes")


def _vp(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _bits(a, shape=None):
    """0/1 array of any integer/bool/float dtype -> contiguous uint8."""
    a = np.asarray(a)
    if a.dtype != np.uint8:
        a = (a.astype(np.int64) & 1).astype(np.uint8) if a.dtype.kind in "iub" else (a != 0).astype(np.uint8)
    a = np.ascontiguousarray(a)
    if shape is not None and a.shape != shape:
        raise ValueError("expected shape %s, got %s" % (shape, a.shape))
    return a


def to_dense(H):
    """Accepts what the reference's callers pass: ndarray of any dtype/order or a scipy.sparse matrix
    (studies/studyComplete.py:84)."""
    if hasattr(H, "toarray") and not isinstance(H, np.ndarray):
        return np.asarray(H.toarray())
    return H if isinstance(H, np.ndarray) else np.asarray(H)


class Code:
    def __init__(self, H, L=None, schedule=(_graph.SEQ, _graph.SEQ), distance=None):
        H = to_dense(H)
        g = _graph.build_graph(H, *schedule)
        self.m, self.n, self.E = g["m"], g["n"], g["E"]
        self.schedule = tuple(schedule)
        self.distance = None if distance is None else int(distance)
        self._g = g
        self.L = None if L is None else _bits(L)
        self.k = 0 if self.L is None else self.L.shape[0]
        if self.L is not None and self.L.shape[1] != self.n:
            raise ValueError("L must have n columns")
        self._h = ctypes.c_void_p()
        lib = _lib.lib()
        _lib.check(lib.qldpc_code_create(self.m, self.n, _vp(g["row_ptr"]), _vp(g["col_idx"]), _vp(g["var_ptr"]),
                                         _vp(g["var_edge0"]), _vp(g["var_edge1"]), self.k, _vp(self.L),
                                         ctypes.byref(self._h)), "qldpc_code_create")
        self.words_m = lib.qldpc_words_m(self._h)
        self.words_n = lib.qldpc_words_n(self._h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().qldpc_code_destroy(h)
            except Exception:
                pass
            self._h = ctypes.c_void_p()

    @property
    def handle(self):
        return self._h

    # ------------------------------------------------------------------------------------------
    @staticmethod
    def config(variant="min_sum", max_iter=50, alpha=1.0, damping=1.0, clip=20.0, precision=32, staged=False,
               lanes_per_shot=0, refill_min=0):
        cfg = _lib.BPConfig()
        cfg.variant = VARIANTS[variant] if isinstance(variant, str) else int(variant)
        cfg.precision = int(precision)
        cfg.max_iter = int(max_iter)
        cfg.staged = int(staged)          # 0 auto | 1 (True) HBM-staged thread-per-shot | 2 thread-per-shot | 3 warp-per-shot | 4 CTA-per-shot | 5 CTA-per-shot, staged
        cfg.lanes_per_shot = int(lanes_per_shot)
        cfg.refill_min = int(refill_min)
        cfg.alpha, cfg.damping, cfg.clip = float(alpha), float(damping), float(clip)
        return cfg

    def _prior(self, prior):
        p = np.ascontiguousarray(np.asarray(prior, dtype=np.float64).reshape(-1))
        if p.size == 1:
            p = np.full(self.n, float(p[0]))
        if p.size != self.n:
            raise ValueError("prior must have n entries")
        return p

    def geometry(self, cfg):
        a, b, c = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
        _lib.check(_lib.lib().qldpc_bp_geometry(self._h, ctypes.byref(cfg), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        kind = c.value
        return dict(shots_per_cta=a.value, smem_bytes=b.value, staged=(kind == 1),
                    lanes_per_shot=(0 if kind in (133, 134) else kind - 100 if kind >= 100 else 1),
                    kernel=('hbm_staged' if kind == 1 else 'cta_staged' if kind == 134 else 'cta_per_shot' if kind == 133
                            else 'warp_per_shot' if kind == 132 else 'tiled' if kind >= 100 else 'thread_per_shot'))

    def tune_warp_layout(self, steps=0):
        """Lane labelling of the warp-per-shot kernels (results never depend on it).  steps = 0: report; steps < 0: install
        the natural labelling; steps > 0: rebuild the constructed (conflict-free) one with that search budget.
        -> dict(natural, current, floor): modelled scatter + gather wavefronts per shot-iteration."""
        cost = (ctypes.c_int32 * 3)()
        _lib.check(_lib.lib().qldpc_warp_layout_tune(self._h, int(steps), cost))
        return dict(natural=cost[0], current=cost[1], floor=cost[2])

    def tiled_conflict_model(self, lanes_per_shot=8):
        """(before, after): modelled shared-memory wavefronts per shot-iteration of the tiled kernel's variable pass with
        natural / optimised variable order."""
        a, b = ctypes.c_double(), ctypes.c_double()
        _lib.check(_lib.lib().qldpc_tiled_conflict_model(self._h, int(lanes_per_shot), ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    # ---- host-array API (reference dtypes) -------------------------------------------------------
    def bp_decode_batch(self, syndromes, prior, variant="min_sum", max_iter=50, alpha=1.0, damping=1.0, clip=20.0,
                        precision=32, want_llr=True, staged=False, lanes_per_shot=0, refill_min=0):
        """-> (hard int8 (B,n), converged bool (B,), llr float64 (B,n) | None, iters int32 (B,))"""
        synd = _bits(syndromes)
        if synd.ndim != 2 or synd.shape[1] != self.m:
            raise ValueError("syndromes must be (B, m)")
        B = synd.shape[0]
        cfg = self.config(variant, max_iter, alpha, damping, clip, precision, staged, lanes_per_shot, refill_min)
        p = self._prior(prior)
        hard = np.zeros((B, self.n), np.int8)
        conv = np.zeros(B, np.uint8)
        iters = np.zeros(B, np.int32)
        llr = np.zeros((B, self.n), np.float64) if want_llr else None
        if B:
            _lib.check(_lib.lib().qldpc_bp_decode_host(self._h, ctypes.byref(cfg), _vp(p), B, _vp(synd), _vp(hard), _vp(conv),
                                                      _vp(iters), _vp(llr)), "qldpc_bp_decode_host")
        return hard, conv.astype(bool), llr, iters

    def bp_messages_batch(self, syndromes, prior, variant="min_sum", max_iter=50, alpha=1.0, damping=1.0, clip=20.0,
                          dump_iter=0):
        """Check-to-variable messages of iteration `dump_iter` as dense (B, m, n) float64 arrays: the
        alpha_estimation=True return of the reference (rework/decoding.py:58-59, :168-169)."""
        synd = _bits(syndromes)
        B = synd.shape[0]
        cfg = self.config(variant, max_iter, alpha, damping, clip, 64)
        p = self._prior(prior)
        r = np.zeros((B, self.E), np.float64)
        if B:
            _lib.check(_lib.lib().qldpc_bp_messages_host(self._h, ctypes.byref(cfg), _vp(p), B, _vp(synd), int(dump_iter), _vp(r)),
                       "qldpc_bp_messages_host")
        rows = np.repeat(np.arange(self.m), np.diff(self._g["row_ptr"]))
        out = np.zeros((B, self.m, self.n), np.float64)
        out[:, rows, self._g["col_idx"]] = r
        return out

    def osd_decode_batch(self, syndromes, llr, hard, order=0, max_combinations=None):
        """-> int64 (B,n): performOSD (order 0) / performOSD_enhanced (order > 0) per shot."""
        synd = _bits(syndromes)
        hard = _bits(hard)
        llr = np.ascontiguousarray(llr, dtype=np.float64)
        B = synd.shape[0]
        if synd.shape != (B, self.m) or hard.shape != (B, self.n) or llr.shape != (B, self.n):
            raise ValueError("shape mismatch")
        out = np.zeros((B, self.n), np.uint8)
        if B:
            _lib.check(_lib.lib().qldpc_osd_decode_host(self._h, B, _vp(synd), _vp(llr), _vp(hard), int(order),
                                                       int(max_combinations) if max_combinations else 0, _vp(out)),
                       "qldpc_osd_decode_host")
        return out.astype(np.int64)

    def bposd_decode_batch(self, syndromes, prior, variant="min_sum", max_iter=50, alpha=1.0, damping=1.0, clip=20.0,
                           precision=32, osd_order=0, staged=False, out=None, lanes_per_shot=0, refill_min=0):
        """BP, then OSD on the BP failures.  -> (corr uint8 (B,n), converged bool (B,), iters int32 (B,))"""
        synd = _bits(syndromes)
        B = synd.shape[0]
        if synd.shape != (B, self.m):
            raise ValueError("syndromes must be (B, m)")
        cfg = self.config(variant, max_iter, alpha, damping, clip, precision, staged, lanes_per_shot, refill_min)
        p = self._prior(prior)
        corr = np.zeros((B, self.n), np.uint8) if out is None else out
        conv = np.zeros(B, np.uint8)
        iters = np.zeros(B, np.int32)
        if B:
            _lib.check(_lib.lib().qldpc_bposd_decode_host(self._h, ctypes.byref(cfg), _vp(p), B, _vp(synd), int(osd_order),
                                                         _vp(corr), _vp(conv), _vp(iters)), "qldpc_bposd_decode_host")
        return corr, conv.astype(bool), iters

    def host_transfer_stats(self):
        """-> dict(h2d_bytes, d2h_bytes: cumulative bytes moved by bposd_decode_batch on this handle; host_pack: -1 not decided,
        0 the byte rows cross PCIe and are packed on the device, 1 host threads pack them, 2 chunk by chunk whichever side is
        free; host_pack_rate: shots/s of the thread pool; chunks_host / chunks_device: chunks that went to either side)"""
        a, b, c, d = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int32(), ctypes.c_double()
        e, f = ctypes.c_uint64(), ctypes.c_uint64()
        _lib.check(_lib.lib().qldpc_host_transfer_stats(self._h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d),
                                                        ctypes.byref(e), ctypes.byref(f)))
        return dict(h2d_bytes=a.value, d2h_bytes=b.value, host_pack=c.value, host_pack_rate=d.value, chunks_host=e.value, chunks_device=f.value)

    def set_host_pack(self, mode):
        """0 / 1 / 2: pack the uint8 rows of bposd_decode_batch on the device / with host threads / chunk by chunk on whichever
        side is free; -1: measure at the next call"""
        _lib.check(_lib.lib().qldpc_set_host_pack(self._h, int(mode)))

    def check_batch(self, errors, corrections, syndromes, converged=None, iters=None, distance=None):
        """-> dict(logical bool (B,), valid bool (B,), degenerate bool (B,), weight int32 (B,), counters dict)"""
        err, cor, syn = _bits(errors), _bits(corrections), _bits(syndromes)
        B = err.shape[0]
        conv = None if converged is None else np.ascontiguousarray(converged, dtype=np.uint8)
        it = None if iters is None else np.ascontiguousarray(iters, dtype=np.int32)
        flags = np.zeros(B, np.uint8)
        weight = np.zeros(B, np.int32)
        counters = np.zeros(_lib.NUM_COUNTERS, np.uint64)
        d = self.distance if distance is None else distance
        if B:
            _lib.check(_lib.lib().qldpc_check_host(self._h, B, _vp(err), _vp(cor), _vp(syn), _vp(conv), _vp(it),
                                                  int(d or 0), _vp(flags), _vp(weight), _vp(counters)), "qldpc_check_host")
        return dict(logical=(flags & 1).astype(bool), valid=(flags & 2).astype(bool), degenerate=(flags & 4).astype(bool),
                    weight=weight, counters=dict(zip(_lib.COUNTER_NAMES, (int(x) for x in counters))))

    def syndromes(self, errors):
        """synd = errors @ H.T % 2 on the device.  errors (B, n) 0/1 -> int8 (B, m)."""
        err = _bits(errors)
        if err.ndim != 2 or err.shape[1] != self.n:
            raise ValueError("errors must be (B, n)")
        syn = np.zeros((err.shape[0], self.m), np.uint8)
        if err.shape[0]:
            _lib.check(_lib.lib().qldpc_syndrome_host(self._h, err.shape[0], _vp(err), _vp(syn)), "qldpc_syndrome_host")
        return syn.view(np.int8)

    def sample(self, p, B, seed=0, first_shot=0, draws=1, meas_p=0.0):
        """Device Philox sampler -> (errors int8 (B,n), syndromes int8 (B,m)).  meas_p > 0: every syndrome bit is flipped
        with that probability (the phenomenological model the reference keeps commented out, paperResults.py:66-68)."""
        err = np.zeros((B, self.n), np.uint8)
        syn = np.zeros((B, self.m), np.uint8)
        if B:
            _lib.check(_lib.lib().qldpc_sample_noisy_host(self._h, float(p), float(meas_p), int(seed), int(first_shot), int(draws), B,
                                                         _vp(err), _vp(syn)), "qldpc_sample_noisy_host")
        return err.view(np.int8), syn.view(np.int8)

    def mc_sweep(self, p, nshots, prior=None, seed=0, first_shot=0, draws=1, variant="min_sum", max_iter=50, alpha=1.0,
                 damping=1.0, clip=20.0, precision=32, osd_order=0, distance=None, staged=False, meas_p=0.0):
        """One Monte-Carlo point entirely on the device.  -> dict of counters.  meas_p: measurement-error rate (see sample)."""
        cfg = self.config(variant, max_iter, alpha, damping, clip, precision, staged)
        pe = p if draws == 1 else 2 * p * (1 - p)
        pr = self._prior(np.log((1 - pe) / pe) if prior is None else prior)
        counters = np.zeros(_lib.NUM_COUNTERS, np.uint64)
        d = self.distance if distance is None else distance
        _lib.check(_lib.lib().qldpc_mc_sweep_noisy(self._h, ctypes.byref(cfg), _vp(pr), float(p), float(meas_p), int(seed),
                                                   int(first_shot), int(nshots), int(draws), int(osd_order), int(d or 0),
                                                   _vp(counters)), "qldpc_mc_sweep_noisy")
        return dict(zip(_lib.COUNTER_NAMES, (int(x) for x in counters)))


    def alpha_counts(self, errors=None, p=None, nshots=0, seed=0, first_shot=0):
        """The four edge counts behind estimate_alpha_from_code (rework/Alvarado.py:10-66), computed on the device:
        counts[bit][s] = number of Tanner-graph edges whose variable has error bit `bit` and whose check has syndrome bit s.
        errors (B, n): the caller's errors; or errors=None: `nshots` device-sampled shots at rate p."""
        counts = np.zeros(4, np.uint64)
        if errors is not None:
            err = _bits(errors)
            if err.ndim != 2 or err.shape[1] != self.n:
                raise ValueError("errors must be (B, n)")
            B, ep = err.shape[0], _vp(err)
        else:
            B, ep = int(nshots), None
        if B:
            _lib.check(_lib.lib().qldpc_alpha_counts(self._h, B, ep, float(p or 0.0), int(seed), int(first_shot), _vp(counts)),
                       "qldpc_alpha_counts")
        return counts.reshape(2, 2).astype(np.int64)

    def llr_histograms(self, p, nshots, lo=-30.0, hi=30.0, nbins=120, prior=None, seed=0, first_shot=0, draws=1,
                       variant="min_sum", max_iter=50, alpha=1.0, damping=1.0, clip=20.0, precision=32):
        """Posterior-LLR histograms computed on the device (no B*n floats returned).
        -> dict(edges float64 (nbins+1,), true_0, true_1, bp_failed_shots uint64 (nbins,), n_bp_failed int)"""
        cfg = self.config(variant, max_iter, alpha, damping, clip, precision)
        pe = p if draws == 1 else 2 * p * (1 - p)
        pr = self._prior(np.log((1 - pe) / pe) if prior is None else prior)
        hist = np.zeros((3, int(nbins)), np.uint64)
        nf = ctypes.c_uint64(0)
        _lib.check(_lib.lib().qldpc_bp_llr_histogram(self._h, ctypes.byref(cfg), _vp(pr), float(p), int(seed), int(first_shot),
                                                     int(nshots), int(draws), float(lo), float(hi), int(nbins), _vp(hist),
                                                     ctypes.byref(nf)), "qldpc_bp_llr_histogram")
        return dict(edges=np.linspace(lo, hi, int(nbins) + 1), true_0=hist[0], true_1=hist[1], bp_failed_shots=hist[2],
                    n_bp_failed=int(nf.value))


# ------------------------------------------------------------------------------------------------
# codes/*.npz and the handle cache used by the reference-named wrappers
# ------------------------------------------------------------------------------------------------
def load_code(name, side="x", directory=None, schedule=None):
    """Loads `codes/<name>.npz` (keys Hx, Hz, Lx, Lz, distance; e.g. name = '[[144, 12, 12]]') and returns a
    Code for H<side> with L<side>, as every reference driver pairs them (paperResults.py:34-39)."""
    d = np.load(os.path.join(directory or DATA_CODES, name + ".npz"))
    H = d["H" + side]
    L = d["L" + side] if ("L" + side) in d.files else None
    dist = int(d["distance"]) if "distance" in d.files else None
    return Code(H, L, schedule or (_graph.SEQ, _graph.SEQ), dist)


_CACHE = {}
_CACHE_MAX = 16


def cached_code(H, variant_key, schedule=None):
    """Code handle for the H object a reference-style call passes, keyed by content, memory layout
    (it selects the float summation order) and variant family."""
    Hd = to_dense(H)
    sched = tuple(schedule) if schedule is not None else _graph.reference_schedule(Hd, variant_key)
    Hb = np.ascontiguousarray(Hd != 0)
    key = (Hb.shape, hashlib.blake2b(Hb.tobytes(), digest_size=16).digest(), sched)
    c = _CACHE.get(key)
    if c is None:
        if len(_CACHE) >= _CACHE_MAX:
            _CACHE.pop(next(iter(_CACHE)))
        c = Code(Hd, None, sched)
        _CACHE[key] = c
    return c
