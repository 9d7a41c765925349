"""TEST INFRASTRUCTURE ONLY -- import the UNMODIFIED reference environments here.

The reference (/root/reference, read-only, never present on the GPU box) imports
plotting / spreadsheet / CPLEX packages that this image does not have.  This
module installs inert stand-ins for the plotting and spreadsheet packages, puts
the docplex shim (oracle/refshim/docplex) on sys.path and returns the reference's
environment classes.  It is used only by oracle/make_golden.py (fixture
generation, run in the build container) and by tests that are skipped when
/root/reference is absent.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("FJSP_REFERENCE_ROOT", "/root/reference")


class _Anything(types.ModuleType):
    """A module whose every attribute is a do-nothing callable/namespace."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        obj = _Blank(name)
        setattr(self, name, obj)
        return obj


class _Blank:
    def __init__(self, name="blank"):
        self._name = name

    def __call__(self, *a, **k):
        return _Blank(self._name)

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Blank(name)

    def __setitem__(self, k, v):
        pass

    def __getitem__(self, k):
        return _Blank()

    def __iter__(self):
        return iter(())


_STUBS = [
    "matplotlib", "matplotlib.pyplot", "matplotlib.font_manager", "matplotlib.ticker",
    "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.axes3d",
    "openpyxl", "openpyxl.styles", "visdom",
]


def install():
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
            continue
        except Exception:
            pass
        mod = _Anything(name)
        mod.__path__ = []
        sys.modules[name] = mod
        if "." in name:
            parent, child = name.rsplit(".", 1)
            setattr(sys.modules[parent], child, mod)
    here = os.path.dirname(os.path.abspath(__file__))
    if here not in sys.path:
        sys.path.insert(0, here)  # docplex shim
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "environments"))


def load(name):
    """name in {'SO_DFJSP','MO_DFJSP','MO_DFJSP_breakdown','SO_FJSSP'} -> env class."""
    install()
    mod = importlib.import_module("environments." + name)
    cls = {"SO_DFJSP": "SO_DFJSP_Environment", "MO_DFJSP": "MO_DFJSP_Environment",
           "MO_DFJSP_breakdown": "MO_DFJSP_Environment", "SO_FJSSP": "SO_FJSSP_Environment"}[name]
    return getattr(mod, cls)
