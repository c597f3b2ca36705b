"""Import the UNMODIFIED reference (michelebanfi/qLDPC) from /root/reference.

Only usable in the build container (the GPU box has no /root/reference); used by
tools/make_golden.py and tools/vendor_codes.py to generate committed fixtures.
Nothing under qldpc_b200/, tests/ (at run time), bench.py or __graft_entry__.py
imports this module.

Quirks handled:
  * decoding/beliefPropagation.py:4 hard-imports drawUtils -> matplotlib (absent):
    a no-op stub module named `drawUtils` is inserted into sys.modules.
  * rework/decoding.py collides with the `decoding/` package name: loaded by path.
"""
import importlib.util
import os
import sys
import types

REF = os.environ.get("QLDPC_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF, "decoding"))


def load():
    """Returns a namespace with the reference's hot-path modules."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REF)
    if "drawUtils" not in sys.modules:
        stub = types.ModuleType("drawUtils")
        stub.plotGraph = lambda *a, **k: None
        stub.plotMatrix = lambda *a, **k: None
        sys.modules["drawUtils"] = stub

    def by_path(name, rel):
        spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    ns = types.SimpleNamespace()
    ns.bp = by_path("_ref_beliefPropagation", "decoding/beliefPropagation.py")
    ns.osd = by_path("_ref_OSD", "decoding/OSD.py")
    ns.osd_enh = by_path("_ref_OSD_enhanced", "decoding/OSD_enhanced.py")
    ns.rework = by_path("_ref_rework_decoding", "rework/decoding.py")
    ns.spacetime = by_path("_ref_spaceTime", "spaceTime.py")
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        ns.bp_gpu = by_path("_ref_beliefPropagationGPU", "decoding/beliefPropagationGPU.py")
    ns.root = REF
    return ns
