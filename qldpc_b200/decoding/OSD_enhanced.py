"""Drop-in for the reference's decoding/OSD_enhanced.py (OSD-w).

OSD-0 runs in osd0_kernel; when its solution misses the syndrome and order > 0 the combination
sweep runs in osdw_kernel (csrc/osdw_kernel.cuh) with the reference's exact enumeration, metric and
selection rule.  Same stable-sort contract as OSD.performOSD.
"""
from .._single import osd_single


def performOSD_enhanced(H, syndrome, llr, hard, order=0, max_combinations=None):
    """Reference: decoding/OSD_enhanced.py:5-131.  Returns the corrected error vector, int64[n]."""
    return osd_single(H, syndrome, llr, hard, order=order, max_combinations=max_combinations)
