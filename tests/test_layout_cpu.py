"""Host logic of the warp- / CTA-per-shot BP kernels: the lane labelling (qldpc_b200/csrc/bp_warp_layout.h) is plain C++;
it is compiled with g++ and checked here without a GPU: conflict-free for every code of the reference and for the full
space-time matrix, every edge in exactly one slot, scatter targets unique and in the right plane, padding confined to the
+inf / dump rows."""
import os
import subprocess
import sys

import numpy as np
import pytest
from scipy.sparse import csr_matrix

from conftest import ROOT, load_code_file


@pytest.fixture(scope="module")
def checker(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("layout") / "layout_check")
    src = os.path.join(ROOT, "tests", "cpp", "layout_check.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", exe, src], check=True, env={k: v for k, v in os.environ.items() if k not in ("CC", "CXX")})
    return exe


def _dump(H, path):
    m, n = H.shape
    Hs = csr_matrix(np.asarray(H) != 0)
    Hs.sort_indices()
    rp, ci = Hs.indptr, Hs.indices
    ec = np.repeat(np.arange(m), np.diff(rp))
    order = np.lexsort((ec, ci))                       # edges per variable, ascending check
    vp = np.concatenate([[0], np.cumsum(np.bincount(ci, minlength=n))])
    with open(path, "w") as f:
        f.write(f"{m} {n}\n")
        for arr in (rp, ci, vp, order, ec):
            f.write(" ".join(map(str, arr)) + "\n")


@pytest.mark.parametrize("stem,args", [("[[72, 12, 6]]", (6, 0, 0)), ("[[90, 8, 10]]", (6, 0, 0)), ("[[108, 8, 10]]", (6, 0, 0)),
                                       ("[[144, 12, 12]]", (6, 0, 0)), ("[[288, 12, 18]]", (6, 0, 0)), ("spacetime", (8, 36, 84))])
def test_lane_labelling_is_conflict_free_and_consistent(checker, tmp_path, stem, args):
    if stem == "spacetime":
        sys.path.insert(0, ROOT)
        from qldpc_b200.spaceTime import spaceTimeMatrix
        H = spaceTimeMatrix(load_code_file("[[144, 12, 12]]")[0], 12)
    else:
        H = load_code_file(stem)[0]
    g = str(tmp_path / "g.txt")
    _dump(H, g)
    r = subprocess.run([checker, g] + [str(a) for a in args], capture_output=True, text=True, timeout=300)
    assert "LAYOUT-OK" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("stem", ["[[72, 12, 6]]", "[[90, 8, 10]]", "[[108, 8, 10]]", "[[144, 12, 12]]", "[[288, 12, 18]]"])
def test_half_warp_labelling_for_64bit_words(checker, tmp_path, stem):
    """float64 kernel: 64-bit shared-memory accesses are served per half-warp; the labelling built with 16-lane conflict
    domains puts the 16 lanes of every half in 16 different bank pairs."""
    g = str(tmp_path / "g.txt")
    _dump(load_code_file(stem)[0], g)
    r = subprocess.run([checker, g, "6", "0", "0", "16"], capture_output=True, text=True, timeout=300)
    assert "LAYOUT-OK" in r.stdout, r.stdout + r.stderr


def _random_regular(m, n, cw, rw, seed):
    """Random bipartite graph with column weight cw and row weight rw (configuration model, no double edges)."""
    rng = np.random.default_rng(seed)
    assert n * cw == m * rw
    while True:
        stubs = np.repeat(np.arange(m), rw)
        rng.shuffle(stubs)
        H = np.zeros((m, n), np.int64)
        ok = True
        for v in range(n):
            cs = stubs[v * cw:(v + 1) * cw]
            if len(set(cs)) < cw:
                ok = False
                break
            H[cs, v] = 1
        if ok:
            return H


@pytest.mark.parametrize("m,n,seed", [(72, 144, 1), (54, 108, 2), (144, 288, 3)])
def test_lane_labelling_on_unstructured_graphs(checker, tmp_path, m, n, seed):
    """The construction does not rely on the bivariate-bicycle structure: random (3, 6)-regular graphs of the same sizes get
    a conflict-free labelling too."""
    H = _random_regular(m, n, 3, 6, seed)
    g = str(tmp_path / "g.txt")
    _dump(H, g)
    r = subprocess.run([checker, g, "6", "0", "0"], capture_output=True, text=True, timeout=300)
    assert "LAYOUT-OK" in r.stdout, r.stdout + r.stderr
