"""Module swap for the reference's own scripts: `install_as_reference_modules()` registers this package under the
import names the reference uses, so that its UNMODIFIED drivers (main.py, paperResults*.py, BP_per_Iteration.py,
rework/main.py, rework/Alvarado.py, studies/studyTT.py) decode on the GPU:

    import qldpc_b200.compat; qldpc_b200.compat.install_as_reference_modules()
    exec(open("paperResults.py").read())          # `from decoding.beliefPropagation import ...` now binds the CUDA path

`decoding` is two things in the reference: the package decoding/ (decoding.beliefPropagation, decoding.OSD,
decoding.OSD_enhanced, decoding.beliefPropagationGPU, decoding.beliefPropagationJAX) for scripts run from the repository
root, and the module rework/decoding.py for scripts run from rework/ (`from decoding import performMinSum_Symmetric`).
The alias serves both: a package whose sub-modules are the former and whose top-level names are the latter (the 4-tuple
performBeliefPropagationFast of rework/decoding.py:77 lives at the top level, the 3-tuple one of
decoding/beliefPropagation.py:88 in the sub-module -- different attribute paths, as in the reference).
`spaceTime` maps to qldpc_b200.spaceTime, `Alvarado` to qldpc_b200.rework.Alvarado; `drawUtils` (matplotlib plotting, out
of scope) gets no-op stand-ins unless the real module is importable.
"""
import importlib
import sys
import types

_SUBMODULES = ("beliefPropagation", "OSD", "OSD_enhanced", "beliefPropagationGPU", "beliefPropagationJAX")
_INSTALLED = {}


def install_as_reference_modules(plot_stubs=True):
    """Registers the aliases in sys.modules (idempotent); returns the names installed."""
    from . import decoding as _dec_pkg
    from .rework import decoding as _rework_dec
    from .rework import Alvarado as _alvarado
    from . import spaceTime as _space_time

    alias = types.ModuleType("decoding")
    alias.__doc__ = "qldpc_b200 alias of the reference's decoding/ package and rework/decoding.py"
    alias.__path__ = list(_dec_pkg.__path__)             # a package: `import decoding.OSD` resolves through sys.modules below
    for name in dir(_rework_dec):
        if name.startswith("perform"):
            setattr(alias, name, getattr(_rework_dec, name))
    mods = {"decoding": alias, "spaceTime": _space_time, "Alvarado": _alvarado}
    for sub in _SUBMODULES:
        m = importlib.import_module("qldpc_b200.decoding." + sub)
        setattr(alias, sub, m)
        mods["decoding." + sub] = m
    if plot_stubs:
        try:
            importlib.import_module("matplotlib")
            have_mpl = True
        except Exception:
            have_mpl = False
        if not have_mpl or "drawUtils" not in sys.modules:
            stub = types.ModuleType("drawUtils")
            stub.plotGraph = lambda *a, **k: None
            stub.plotMatrix = lambda *a, **k: None
            stub.__doc__ = "no-op stand-ins for the reference's matplotlib helpers (out of scope)"
            mods["drawUtils"] = stub
    for k, v in mods.items():
        _INSTALLED.setdefault(k, sys.modules.get(k))
        sys.modules[k] = v
    return sorted(mods)


def uninstall_reference_modules():
    """Restores sys.modules to what it held before install_as_reference_modules()."""
    for k, old in list(_INSTALLED.items()):
        if old is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = old
        del _INSTALLED[k]
