"""Re-pack the reference's code matrices (codes/*.npz) into
qldpc_b200/data/codes/ so that tests and bench.py can run on the GPU box, where
/root/reference does not exist.

Same file names, same keys (Hx, Hz, Lx, Lz, distance), same dtypes and the same
memory order (Hx is Fortran-ordered in the BB files; that matters for the float
summation schedule) -- only zip-deflated (1.05 MB -> ~30 KB).
These are data (matrices), not source.
"""
import glob
import os
import numpy as np

SRC = "/root/reference/codes"
DST = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qldpc_b200", "data", "codes")

if __name__ == "__main__":
    os.makedirs(DST, exist_ok=True)
    for f in sorted(glob.glob(os.path.join(SRC, "*.npz"))):
        d = np.load(f)
        arrs = {k: d[k] for k in d.files}
        out = os.path.join(DST, os.path.basename(f))
        np.savez_compressed(out, **arrs)
        chk = np.load(out)
        for k in d.files:
            assert chk[k].dtype == d[k].dtype and np.array_equal(chk[k], d[k])
            if d[k].ndim == 2:
                assert chk[k].flags["F_CONTIGUOUS"] == d[k].flags["F_CONTIGUOUS"], (f, k)
        print(out, os.path.getsize(out))
