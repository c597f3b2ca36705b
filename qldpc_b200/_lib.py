"""ctypes binding of libqldpc_b200.so (include/qldpc_b200.h).  Fails loudly: there is no CPU
fallback anywhere in this package."""
import ctypes
import os

from . import build as _build

_LIB = None

c_i32, c_i64, c_u64, c_dbl, c_vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64, ctypes.c_double, ctypes.c_void_p

QLDPC_OK = 0
MIN_SUM, SUM_PRODUCT, SUM_PRODUCT_SYM = 0, 1, 2
LLR_NONE, LLR_FAILED, LLR_ALL = 0, 1, 2
NUM_COUNTERS = 16
COUNTER_NAMES = ["shots", "bp_failed", "logical", "logical_and_osd", "degenerate", "miscorrected", "incorrectable",
                 "invalid", "iter_sum", "logical_and_bp_converged", "residual_weight", "error_weight"]


class BPConfig(ctypes.Structure):
    _fields_ = [("variant", c_i32), ("precision", c_i32), ("max_iter", c_i32), ("staged", c_i32),
                ("lanes_per_shot", c_i32), ("refill_min", c_i32),
                ("alpha", c_dbl), ("damping", c_dbl), ("clip", c_dbl)]


class QldpcError(RuntimeError):
    pass


def lib():
    """Loads the CUDA library; raises if it is not built (run `python -m qldpc_b200.build`)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not os.path.exists(path):
        raise QldpcError("libqldpc_b200.so is not built: run `python -m qldpc_b200.build` "
                         "(needs nvcc; there is no CPU fallback)")
    L = ctypes.CDLL(path)
    L.qldpc_last_error.restype = ctypes.c_char_p
    P = ctypes.POINTER
    sigs = {
        "qldpc_version": ([], ctypes.c_int),
        "qldpc_device_count": ([], ctypes.c_int),
        "qldpc_code_create": ([c_i32, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, P(c_vp)], ctypes.c_int),
        "qldpc_code_destroy": ([c_vp], None),
        "qldpc_bp_geometry": ([c_vp, P(BPConfig), P(c_i32), P(c_i32), P(c_i32)], ctypes.c_int),
        "qldpc_tiled_conflict_model": ([c_vp, c_i32, P(c_dbl), P(c_dbl)], ctypes.c_int),
        "qldpc_warp_layout_tune": ([c_vp, c_i64, P(c_i32)], ctypes.c_int),
        "qldpc_words_m": ([c_vp], ctypes.c_int),
        "qldpc_words_n": ([c_vp], ctypes.c_int),
        "qldpc_bp_decode_host": ([c_vp, P(BPConfig), c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_bp_messages_host": ([c_vp, P(BPConfig), c_vp, c_i64, c_vp, c_i32, c_vp], ctypes.c_int),
        "qldpc_osd_decode_host": ([c_vp, c_i64, c_vp, c_vp, c_vp, c_i32, c_i64, c_vp], ctypes.c_int),
        "qldpc_bposd_decode_host": ([c_vp, P(BPConfig), c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_bposd_decode_host_packed": ([c_vp, P(BPConfig), c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_host_transfer_stats": ([c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_set_host_pack": ([c_vp, c_i32], ctypes.c_int),
        "qldpc_check_host": ([c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_syndrome_host": ([c_vp, c_i64, c_vp, c_vp], ctypes.c_int),
        "qldpc_syndrome_dev": ([c_vp, c_i64, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_sample_host": ([c_vp, c_dbl, c_u64, c_u64, c_i32, c_i64, c_vp, c_vp], ctypes.c_int),
        "qldpc_mc_sweep": ([c_vp, P(BPConfig), c_vp, c_dbl, c_u64, c_u64, c_i64, c_i32, c_i32, c_i32, c_vp], ctypes.c_int),
        "qldpc_mc_sweep_noisy": ([c_vp, P(BPConfig), c_vp, c_dbl, c_dbl, c_u64, c_u64, c_i64, c_i32, c_i32, c_i32, c_vp], ctypes.c_int),
        "qldpc_sample_noisy_host": ([c_vp, c_dbl, c_dbl, c_u64, c_u64, c_i32, c_i64, c_vp, c_vp], ctypes.c_int),
        "qldpc_osdw_decode_dev": ([c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_i32, c_i64, c_vp], ctypes.c_int),
        "qldpc_alpha_counts": ([c_vp, c_i64, c_vp, c_dbl, c_u64, c_u64, c_vp], ctypes.c_int),
        "qldpc_measurement_noise_dev": ([c_vp, c_dbl, c_u64, c_u64, c_i64, c_vp, c_vp], ctypes.c_int),
        "qldpc_bp_llr_histogram": ([c_vp, P(BPConfig), c_vp, c_dbl, c_u64, c_u64, c_i64, c_i32, c_dbl, c_dbl, c_i32, c_vp, c_vp], ctypes.c_int),
        "qldpc_bp_decode_dev": ([c_vp, P(BPConfig), c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_osd_decode_dev": ([c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_check_dev": ([c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_sample_dev": ([c_vp, c_dbl, c_u64, c_u64, c_i32, c_i64, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_bposd_decode_dev": ([c_vp, P(BPConfig), c_vp, c_i64, c_vp, c_i32, c_vp, c_vp, c_vp, c_vp, c_vp], ctypes.c_int),
        "qldpc_pack_bits_dev": ([c_vp, c_vp, c_i64, c_i32, c_vp], ctypes.c_int),
        "qldpc_unpack_bits_dev": ([c_vp, c_vp, c_i64, c_i32, c_vp], ctypes.c_int),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(L, name)   # AttributeError if the library does not export a declared symbol
        fn.argtypes = argtypes
        fn.restype = restype
    _LIB = L
    return L


EXPORTED = ["qldpc_last_error", "qldpc_version", "qldpc_device_count", "qldpc_code_create", "qldpc_code_destroy",
            "qldpc_bp_geometry", "qldpc_tiled_conflict_model", "qldpc_warp_layout_tune", "qldpc_words_m", "qldpc_words_n", "qldpc_bp_decode_host", "qldpc_bp_messages_host", "qldpc_osd_decode_host",
            "qldpc_bposd_decode_host", "qldpc_bposd_decode_host_packed", "qldpc_host_transfer_stats", "qldpc_set_host_pack", "qldpc_check_host", "qldpc_syndrome_host", "qldpc_syndrome_dev", "qldpc_sample_host", "qldpc_sample_noisy_host", "qldpc_mc_sweep", "qldpc_mc_sweep_noisy", "qldpc_measurement_noise_dev", "qldpc_alpha_counts",
            "qldpc_bp_llr_histogram", "qldpc_bp_decode_dev",
            "qldpc_osd_decode_dev", "qldpc_osdw_decode_dev", "qldpc_check_dev", "qldpc_sample_dev", "qldpc_bposd_decode_dev",
            "qldpc_pack_bits_dev", "qldpc_unpack_bits_dev"]


def check(rc, what=""):
    if rc != QLDPC_OK:
        msg = lib().qldpc_last_error()
        raise QldpcError("%s failed (code %d): %s" % (what or "qldpc call", rc, msg.decode() if msg else "?"))
