"""Vendors the UNMODIFIED hot-path files of the reference into baseline/_ref/ (git-ignored; travels to the GPU box with
the gpurun snapshot, like the built .so files), so that bench.py can time the reference's own Python functions on the GPU
box's host cores (cpu_baseline kind "reference_python") and tests can execute the reference's own driver scripts against the
module swap (qldpc_b200.compat).  Nothing under baseline/_ref/ is ever committed or imported by the product path.

    python tools/vendor_reference.py [/root/reference]

`pip install --target baseline/_ref /root/reference` is not possible: the reference has neither setup.py nor
pyproject.toml ("Directory is not installable"); its files are copied verbatim instead.
"""
import glob
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["main.py", "paperResults.py", "paperResults_GPU.py", "BP_per_Iteration.py", "spaceTime.py", "loadResults.py",
         "decoding/beliefPropagation.py", "decoding/beliefPropagationGPU.py", "decoding/OSD.py", "decoding/OSD_enhanced.py",
         "rework/decoding.py", "rework/Alvarado.py", "rework/main.py", "rework/main_different_orders.py", "studies/studyTT.py"]


def vendor(src="/root/reference"):
    if not os.path.isdir(src):
        return None
    manifest = {}
    files = list(FILES) + [os.path.relpath(p, src) for p in sorted(glob.glob(os.path.join(src, "codes", "*.npz")))]
    for rel in files:
        s, d = os.path.join(src, rel), os.path.join(DEST, rel)
        if not os.path.exists(s):
            continue
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        manifest[rel] = hashlib.sha256(open(s, "rb").read()).hexdigest()
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": src, "sha256": manifest}, f, indent=1, sort_keys=True)
    return DEST


if __name__ == "__main__":
    print(vendor(*(sys.argv[1:2])))
