"""In-tree build of libqldpc_b200.so (hand-written CUDA for sm_100a behind a C ABI).

    python -m qldpc_b200.build [--force]

nvcc cross-compiles without a GPU; the built .so is git-ignored but travels to the GPU box.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_DIR = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libqldpc_b200.so")
NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
    "-Xptxas", "-v",
]


def sources():
    out = [os.path.join(HERE, "..", "include", "qldpc_b200.h")]
    for f in sorted(os.listdir(SRC_DIR)):
        if f.endswith((".cu", ".cuh", ".h")):
            out.append(os.path.join(SRC_DIR, f))
    return out


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(s) > t for s in sources())


def build_library(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    cmd = [nvcc] + NVCC_FLAGS + ["-o", LIB, os.path.join(SRC_DIR, "capi.cu")]
    env = dict(os.environ)
    env.pop("CC", None)   # the image exports a CC that nvcc must not pick up as host compiler
    env.pop("CXX", None)
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(" ".join(cmd) + "\n" + res.stdout)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed (see qldpc_b200/build.log)")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
