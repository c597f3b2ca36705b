"""Runs one of the reference's own driver scripts (vendored UNMODIFIED under baseline/_ref/ by tools/vendor_reference.py)
either on the reference's own modules or on the qldpc_b200 module swap (qldpc_b200.compat), in a scratch directory, and
returns the script's namespace.  Plotting (matplotlib, absent here and out of scope) is replaced by an inert object; the
only edits applied to the script TEXT are the ones the caller passes (trial counts / code lists, to bound the run time)."""
import contextlib
import io
import os
import re
import sys
import types

from conftest import ROOT

REF = os.path.join(ROOT, "baseline", "_ref")


class _Inert:
    """plt.anything(...).anything[...] -> itself; unpacks as a pair (fig, axes = plt.subplots(...))."""
    def __getattr__(self, name): return self
    def __call__(self, *a, **k): return self
    def __getitem__(self, k): return self
    def __iter__(self): return iter((self, self))
    def __enter__(self): return self
    def __exit__(self, *a): return False


def available():
    return os.path.exists(os.path.join(REF, "main.py"))


def _purge(names):
    for k in list(sys.modules):
        if k.split(".")[0] in names:
            del sys.modules[k]


def _snapshot(x):
    import numpy as np
    if isinstance(x, np.ndarray):
        return x.copy()
    if isinstance(x, tuple):
        return tuple(_snapshot(v) for v in x)
    if isinstance(x, list):
        try:
            return np.array(x)
        except Exception:
            return list(x)
    return x


def run_script(rel, swap, edits=(), from_rework=False, workdir=None, record=()):
    """Executes baseline/_ref/<rel> as __main__.  swap=False: the reference's own decoders; swap=True: the CUDA path.
    record: (module, function) names whose calls are logged -- ns["__calls__"][function] = [(args, result), ...] (copies)."""
    src = open(os.path.join(REF, rel)).read()
    for pat, rep in edits:
        src, cnt = re.subn(pat, rep, src)
        assert cnt >= 1, "edit %r did not apply to %s" % (pat, rel)
    names = {"decoding", "spaceTime", "drawUtils", "Alvarado", "matplotlib"}
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in names}
    saved_path = list(sys.path)
    cwd = os.getcwd()
    _purge(names)
    inert = _Inert()
    mpl = types.ModuleType("matplotlib")
    mpl.pyplot = inert
    mpl.use = lambda *a, **k: None
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = inert
    draw = types.ModuleType("drawUtils")
    draw.plotGraph = draw.plotMatrix = lambda *a, **k: None
    sys.modules["drawUtils"] = draw
    wrapped = []
    try:
        os.chdir(workdir)
        for d in ("data", "media", "rework"):
            os.makedirs(d, exist_ok=True)
        if not os.path.exists("codes"):
            os.symlink(os.path.join(REF, "codes"), "codes")
        if swap:
            import qldpc_b200.compat as compat
            compat.install_as_reference_modules()
        else:
            sys.path.insert(0, REF)
            if from_rework:
                sys.path.insert(0, os.path.join(REF, "rework"))     # `decoding` = rework/decoding.py, as when run from rework/
        calls = {}
        for modname, fname in record:
            import importlib
            mod = importlib.import_module(modname)
            orig = getattr(mod, fname)
            wrapped.append((mod, fname, orig))

            def make(orig=orig, key=fname):
                def f(*a, **k):
                    args = tuple(_snapshot(v) for v in a)
                    r = orig(*a, **k)
                    calls.setdefault(key, []).append((args, _snapshot(r)))
                    return r
                return f
            setattr(mod, fname, make())
        ns = {"__name__": "__main__", "__file__": os.path.join(REF, rel), "__calls__": calls}
        out = io.StringIO()
        with contextlib.redirect_stdout(out), contextlib.redirect_stderr(io.StringIO()):
            exec(compile(src, rel, "exec"), ns)
        ns["__stdout__"] = out.getvalue()
        return ns
    finally:
        for mod, fname, orig in wrapped:
            setattr(mod, fname, orig)
        os.chdir(cwd)
        if swap:
            import qldpc_b200.compat as compat
            compat.uninstall_reference_modules()
        _purge(names)
        sys.modules.update(saved)
        sys.path[:] = saved_path
