"""Drop-in for the reference's decoding/OSD.py (OSD-0), running osd0_kernel (csrc/osd_kernel.cuh).

Contract difference, stated in SURVEY.md H1: columns are ordered by a STABLE ascending sort of
|llr| (ties -> lower index); the reference's np.argsort(kind='quicksort') tie order is unspecified.
Unlike the reference (OSD.py:67 raises TypeError on a float64 H) any numeric H is accepted.
"""
from .._single import osd_single


def performOSD(H, syndrome, llr, hard):
    """Reference: decoding/OSD.py:3-28.  Returns the corrected error vector, int64[n]."""
    return osd_single(H, syndrome, llr, hard, order=0)
