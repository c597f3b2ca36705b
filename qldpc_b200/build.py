"""In-tree build of libqldpc_b200.so (hand-written CUDA for sm_100a behind a C ABI).

    python -m qldpc_b200.build [--force] [-v]

One object per csrc/*.cu (a kernel family each), compiled in parallel and linked into one shared library.
nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libqldpc_b200.so")
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def units():
    return sorted(f for f in os.listdir(SRC_DIR) if f.endswith(".cu"))


def headers():
    out = [os.path.join(HERE, "..", "include", "qldpc_b200.h")]
    for f in sorted(os.listdir(SRC_DIR)):
        if f.endswith((".cuh", ".h")):
            out.append(os.path.join(SRC_DIR, f))
    return out


def sources():
    return headers() + [os.path.join(SRC_DIR, f) for f in units()]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def _nvcc():
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return nvcc if os.path.exists(nvcc) else "nvcc"


def _env():
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a CC that nvcc must not pick up as host compiler
    env.pop("CXX", None)
    return env


def _compile(unit, force):
    src = os.path.join(SRC_DIR, unit)
    obj = os.path.join(OBJ_DIR, unit[:-3] + ".o")
    if not force and os.path.exists(obj):
        t = os.path.getmtime(obj)
        if os.path.getmtime(src) <= t and all(os.path.getmtime(h) <= t for h in headers()):
            return obj, 0, "(up to date) %s\n" % unit
    cmd = [_nvcc()] + NVCC_FLAGS + os.environ.get("QLDPC_NVCC_EXTRA", "").split() + ["-c", "-o", obj, src]     # (diagnostic builds)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=_env())
    return obj, res.returncode, " ".join(cmd) + "\n" + res.stdout


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OBJ_DIR, exist_ok=True)
    with ThreadPoolExecutor(max_workers=max(1, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda u: _compile(u, force), units()))
    log = "".join(r[2] for r in results)
    rc = max(r[1] for r in results)
    if rc == 0:
        cmd = [_nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + [r[0] for r in results]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=_env())
        log += " ".join(cmd) + "\n" + res.stdout
        rc = res.returncode
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(log)
    if verbose or rc != 0:
        print(log)
    if rc != 0:
        raise RuntimeError("nvcc failed (see qldpc_b200/build.log)")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
