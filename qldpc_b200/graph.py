"""Host-side Tanner-graph preprocessing: CSR of H, per-variable edge tables, and the ORDER in which
the check-to-variable messages of a variable are added into its posterior.

Why the order matters: the reference computes `np.sum(R, axis=0)` on a dense (m, n)
float64 array (decoding/beliefPropagation.py:129, rework/decoding.py:61,173).  NumPy adds the m rows
sequentially when the array is C-ordered, but with its pairwise-8 scheme when the reduction runs
along the contiguous axis of a Fortran-ordered array -- and `Hx` in every BB codes/*.npz is
Fortran-ordered.  Floating-point addition is not associative, so to reproduce the reference's
float64 results bit for bit the float64 kernel adds in the same order.  For columns of weight <= 3
(every matrix the reference uses) any summation tree is a chain (x + y) + z, so an order table is
enough; the chain is found by PROBING NumPy itself with values that make the association visible,
not by re-deriving its algorithm.
"""
import numpy as np

SEQ, PAIRWISE = "sequential", "pairwise"


def is_fortran_only(H):
    return (isinstance(H, np.ndarray) and H.ndim == 2 and H.flags["F_CONTIGUOUS"] and not H.flags["C_CONTIGUOUS"])


def reference_schedule(H, variant):
    """(order mode of iteration 0, order mode of iterations >= 1) that NumPy uses inside the reference
    function `variant` ('min_sum' | 'sum_product' | 'sum_product_sym' | 'loop') for this H object."""
    if variant == "loop" or not is_fortran_only(H):
        return SEQ, SEQ          # beliefPropagation.py:68 sums a gathered (<8 element) vector sequentially
    if variant == "sum_product":
        return PAIRWISE, PAIRWISE
    # min-sum / symmetric sum-product: `Q_old = Q.copy()` (decoding.py:22,150) is C-ordered, and
    # `damping * Q_new + (1 - damping) * Q_old` (:65,:179) only stays Fortran-ordered when NumPy elides the
    # temporary, i.e. for arrays of at least 256 KiB.
    big = H.shape[0] * H.shape[1] * 8 >= 256 * 1024
    return PAIRWISE, (PAIRWISE if big else SEQ)


def _probe_first_pair(m, rows3):
    """For columns with exactly three non-zero rows (rows3: (ncols, 3) ascending), find which two rows
    NumPy's sum over a contiguous length-m column combines first.  Returns (ncols,) index of the row
    added LAST (0, 1 or 2)."""
    ncols = rows3.shape[0]
    tiny, one = 2.0 ** -53, 1.0
    last = np.full(ncols, -1)
    for z in range(3):                       # hypothesis: rows3[:, z] is added last
        A = np.zeros((m, ncols), order="F")
        for k in range(3):
            A[rows3[:, k], np.arange(ncols)] = one if k == z else tiny
        s = np.sum(A, axis=0)                # reduction along the contiguous axis: pairwise
        # (tiny + tiny) + 1 = 1 + 2^-52, whereas (1 + tiny) + tiny = 1 (ties to even, twice)
        hit = s == one + 2.0 ** -52
        last[hit & (last < 0)] = z
    if (last < 0).any():
        raise RuntimeError("could not determine NumPy's summation order by probing")
    return last


def build_graph(H, mode0=SEQ, mode1=SEQ):
    """Returns a dict of int32 arrays: row_ptr, col_idx (CSR, ascending columns), var_ptr, var_edge0,
    var_edge1 (edge ids of each variable in addition order for iteration 0 / >= 1), plus m, n, E."""
    Hb = np.asarray(H) != 0
    if Hb.ndim != 2:
        raise ValueError("H must be a 2-D matrix")
    m, n = Hb.shape
    rows, cols = np.nonzero(Hb)               # row-major scan: ascending column inside each row
    E = rows.size
    row_ptr = np.zeros(m + 1, np.int32)
    np.cumsum(np.bincount(rows, minlength=m), out=row_ptr[1:])
    # variable-major view: stable sort of the edges by column keeps ascending check order
    by_var = np.argsort(cols, kind="stable").astype(np.int32)
    var_ptr = np.zeros(n + 1, np.int32)
    np.cumsum(np.bincount(cols, minlength=n), out=var_ptr[1:])
    tables = {SEQ: by_var}
    if PAIRWISE in (mode0, mode1):
        deg = np.diff(var_ptr)
        if deg.max() > 3:
            raise NotImplementedError("pairwise (Fortran-order) summation schedule for column weight > 3")
        pw = by_var.copy()
        v3 = np.nonzero(deg == 3)[0]
        if v3.size and m >= 8:                # NumPy sums fewer than 8 rows sequentially
            e3 = by_var[var_ptr[v3][:, None] + np.arange(3)[None, :]]          # (nv3, 3) edge ids, ascending check
            last = _probe_first_pair(m, rows[e3])
            keep = np.array([[1, 2], [0, 2], [0, 1]])[last]                   # the pair added first (ascending)
            order = np.concatenate([np.take_along_axis(e3, keep, 1), np.take_along_axis(e3, last[:, None], 1)], 1)
            pw[var_ptr[v3][:, None] + np.arange(3)[None, :]] = order
        tables[PAIRWISE] = pw
    return dict(m=m, n=n, E=int(E), row_ptr=row_ptr, col_idx=cols.astype(np.int32), var_ptr=var_ptr,
                var_edge0=np.ascontiguousarray(tables[mode0], np.int32),
                var_edge1=np.ascontiguousarray(tables[mode1], np.int32))
