"""CPU tests of the host-side logic and of the C-ABI library's surface (no compute calls)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_code_file
from oracle import oracle as O
from qldpc_b200 import graph as G


def test_library_exports_every_declared_symbol():
    from qldpc_b200 import _lib, build
    build.build_library()
    header = open(os.path.join(ROOT, "include", "qldpc_b200.h")).read()
    declared = set(re.findall(r"\b(qldpc_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    raw = ctypes.CDLL(build.LIB)
    for name in sorted(declared):
        assert hasattr(raw, name), "library does not export %s" % name
    assert declared == set(_lib.EXPORTED)
    L = _lib.lib()           # binds argtypes for every symbol
    assert L.qldpc_version() >= 100


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail loudly."""
    from qldpc_b200 import _lib, Code, QldpcError
    if _lib.lib().qldpc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    H, _ = load_code_file("[[72, 12, 6]]")
    with pytest.raises(QldpcError):
        Code(H)


@pytest.mark.parametrize("stem", ["steane", "[[72, 12, 6]]", "[[90, 8, 10]]", "[[108, 8, 10]]", "[[144, 12, 12]]", "[[288, 12, 18]]"])
def test_graph_tables_and_schedule_probe(stem):
    """graph.py finds NumPy's summation order by probing NumPy; the oracle derives it symbolically.
    Both must describe the same sums."""
    H, _ = load_code_file(stem)
    g = G.build_graph(H, G.PAIRWISE, G.SEQ)
    o = O.Graph(H, O.PAIRWISE, O.SEQ)
    assert np.array_equal(g["row_ptr"], o.row_ptr) and np.array_equal(g["col_idx"], o.col_idx)
    assert np.array_equal(g["var_ptr"], o.var_ptr) and np.array_equal(g["var_edge1"], o.var_edge1)
    for v in range(g["n"]):
        a = g["var_edge0"][g["var_ptr"][v]:g["var_ptr"][v + 1]]
        b = o.var_edge0[o.var_ptr[v]:o.var_ptr[v + 1]]
        if len(a) == 3:       # (x + y) + z: the first two commute
            assert set(a[:2]) == set(b[:2]) and a[2] == b[2]
        else:
            assert sorted(a) == sorted(b)
    # the probed order really reproduces np.sum on a Fortran-ordered array
    rng = np.random.default_rng(0)
    R = np.where(np.asarray(H) != 0, rng.normal(0, 3, H.shape), 0.0)
    Rf = np.asfortranarray(R)
    want = np.sum(Rf, axis=0)
    edge_val = R[np.repeat(np.arange(g["m"]), np.diff(g["row_ptr"])), g["col_idx"]]
    for v in range(g["n"]):
        e = g["var_edge0"][g["var_ptr"][v]:g["var_ptr"][v + 1]]
        s = edge_val[e[0]]
        for x in e[1:]:
            s = s + edge_val[x]
        assert s == want[v]


def test_reference_schedule_rules():
    H, _ = load_code_file("[[144, 12, 12]]")
    assert G.reference_schedule(H, "sum_product") == (G.PAIRWISE, G.PAIRWISE)
    assert G.reference_schedule(H, "min_sum") == (G.PAIRWISE, G.SEQ)
    assert G.reference_schedule(np.ascontiguousarray(H), "min_sum") == (G.SEQ, G.SEQ)
    H288, _ = load_code_file("[[288, 12, 18]]")
    assert G.reference_schedule(H288, "min_sum") == (G.PAIRWISE, G.PAIRWISE)
    assert G.reference_schedule(H, "loop") == (G.SEQ, G.SEQ)
    for v, ov in (("sum_product", O.SUM_PRODUCT), ("min_sum", O.MIN_SUM), ("sum_product_sym", O.SUM_PRODUCT_SYM)):
        for Hx in (H, H288, np.ascontiguousarray(H)):
            assert G.reference_schedule(Hx, v) == O.auto_schedule(Hx, ov)


def test_spacetime_host(spacetime_golden):
    from qldpc_b200.spaceTime import spaceTimeMatrix, spacetimeSyndrome
    d = spacetime_golden
    H, _ = load_code_file("[[72, 12, 6]]")
    Hst = spaceTimeMatrix(H, 3)
    ref = np.zeros(tuple(d["Hst_shape"]))
    ref[d["Hst_rows"], d["Hst_cols"]] = 1.0
    assert Hst.dtype == np.float64 and Hst.flags["C_CONTIGUOUS"] and np.array_equal(Hst, ref)
    assert np.array_equal(spaceTimeMatrix(H, 1), O.space_time_matrix(H, 1))
    np.random.seed(5)
    e, s = spacetimeSyndrome(H, 0.03, 3)
    assert np.array_equal(e, d["seed5_error"]) and np.array_equal(s, d["seed5_syndrome"])


def test_save_results_round_trip_like_load_results(tmp_path):
    """experiments.save_results writes what every reference script writes (np.savez(path, results=dict)); loadResults.py:5-12
    reads it back with np.load(..., allow_pickle=True)["results"].item()."""
    import numpy as np
    from qldpc_b200 import experiments as X
    res = {"[[72, 12, 6]]": {"ler": [0.1, 0.01], "BPs_fault": [0, 0], "degeneracies": [3, 1]}}
    path = str(tmp_path / "BPOSD.npz")
    X.save_results(path, res)
    f = np.load(path, allow_pickle=True)
    assert list(f.keys()) == ["results"]
    assert f["results"].item() == res


@pytest.mark.parametrize("flags", [[], ["-DSPM_EMULATE_RCP"]])
def test_sum_product_math_header_against_50_digit_references(tmp_path, flags):
    """qldpc_b200/csrc/sp_math.cuh (the float64 tanh(q/2) / 2 atanh(clip(x)) of the warp-per-shot sum-product kernels), built
    for the host -- once with the host's division, once with a model of the device's reciprocal sequence (20-bit seed + one
    cubic Newton step) -- against mpmath at 40 digits: relative error <= 5e-16, signs and zeros as tanh / arctanh."""
    import subprocess
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    lib = str(tmp_path / "spm.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", *flags, "-o", lib,
                    os.path.join(ROOT, "tests", "cpp", "sp_math_check.cpp")], check=True,
                   env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    L = ctypes.CDLL(lib)

    def call(f, x):
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        f(x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size))
        return y

    rng = np.random.default_rng(0)
    q = np.concatenate([rng.uniform(-60, 60, 1500), rng.normal(0, 3, 1500), 10.0 ** rng.uniform(-18, 1, 1000) * rng.choice([-1, 1], 1000),
                        [0.0, 1e-300, 80, 100, -200, 0.69314718, 0.34657359, -0.34657359, 37.0, 38.5]])
    t = call(L.sp_tanh_half, q)
    ref = np.array([float(mp.tanh(mp.mpf(float(v)) / 2)) for v in q])
    assert (np.abs(t - ref) <= 5e-16 * np.abs(ref)).all()
    x = np.concatenate([rng.uniform(-1, 1, 2000), 1 - 10.0 ** rng.uniform(-7.5, -0.3, 1000), -(1 - 10.0 ** rng.uniform(-7.5, -0.3, 500)),
                        10.0 ** rng.uniform(-18, -0.5, 1000), [0, 0.2, 0.1999999, 0.2000001, 0.5, 0.9999999, 1.0, -1.0, 0.99999995, 1 / 3, 0.6]])
    r = call(L.sp_2atanh, x)
    ref = np.array([float(2 * mp.atanh(mp.mpf(float(v)))) for v in np.clip(x, -0.9999999, 0.9999999)])
    assert (np.abs(r - ref) <= 5e-16 * np.abs(ref)).all()
    assert r[np.abs(x) >= 0.9999999].tolist() == [2 * np.arctanh(0.9999999) * np.sign(v) for v in x[np.abs(x) >= 0.9999999]] or \
        np.allclose(np.abs(r[np.abs(x) >= 0.9999999]), 2 * np.arctanh(0.9999999), rtol=5e-16)
    z = call(L.sp_tanh_half, [0.0, -0.0, -3.0])
    assert z[0] == 0 and not np.signbit(z[0]) and z[1] == 0 and np.signbit(z[1]) and z[2] < 0
    z = call(L.sp_2atanh, [0.0, -0.0, -0.5])
    assert z[0] == 0 and not np.signbit(z[0]) and z[1] == 0 and np.signbit(z[1]) and z[2] < 0


def test_host_side_bit_packing_matches_numpy(tmp_path):
    """qldpc_b200/csrc/host_pack.h (the host-thread packing of qldpc_bposd_decode_host's uint8 rows), built with g++: bytes ->
    bits equals np.packbits(little) on `byte & 1`, bits -> bytes is its inverse, for row lengths with and without full words,
    aligned and unaligned outputs (non-temporal and plain store paths), several thread counts."""
    import subprocess
    lib = str(tmp_path / "hp.so")
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", lib, os.path.join(ROOT, "tests", "cpp", "host_pack_check.cpp")],
                   check=True, env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    L = ctypes.CDLL(lib)
    rng = np.random.default_rng(0)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    for nbits, threads in ((72, 3), (144, 4), (36, 1), (45, 2), (90, 5), (7, 2), (33, 3), (64, 2), (100, 3), (2592, 4), (864, 2)):
        B = 257
        W = (nbits + 31) // 32
        a = (rng.random((B, nbits)) < 0.3).astype(np.uint8)
        a[::7] |= 2                                              # only bit 0 of a byte counts (as in pack_bits_kernel)
        out = np.zeros((B, W), dtype=np.uint32)
        L.hp_pack(vp(a), vp(out), ctypes.c_longlong(B), nbits, threads)
        bits = a & 1
        pad = np.zeros((B, W * 32), dtype=np.uint8)
        pad[:, :nbits] = bits
        ref = np.packbits(pad.reshape(B, W, 32), axis=2, bitorder="little").view(np.uint32).reshape(B, W)
        assert np.array_equal(out, ref), nbits
        for shift in (0, 1):
            buf = np.full(B * nbits + 64, 7, dtype=np.uint8)
            off = (-buf.ctypes.data) % 16 + shift
            o = buf[off:off + B * nbits]
            L.hp_unpack(vp(ref), vp(o), ctypes.c_longlong(B), nbits, threads)
            assert np.array_equal(o.reshape(B, nbits), bits), (nbits, shift)
            assert (buf[:off] == 7).all() and (buf[off + B * nbits:] == 7).all()
