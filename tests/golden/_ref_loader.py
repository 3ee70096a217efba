"""Import the read-only reference's numerical core in THIS container.

iris / cf_units / cartopy / netCDF4 are absent from the image, so they are
replaced by MagicMock modules before `improver` is imported.
Only array-level functions are used afterwards; nothing under /root/reference
is copied.  This module is used ONLY by make_golden.py (fixture generation)
and by the optional `-m "not gpu"` cross-check tests, which skip when
/root/reference is absent (it does not exist on the GPU box).
"""
import importlib.abc
import importlib.machinery
import os
import sys
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("IMPROVER_REFERENCE", "/root/reference")

_STUB_ROOTS = ("iris", "cf_units", "cftime", "cartopy", "netCDF4", "dask",
               "clize", "sigtools", "stratify")


class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Serve a MagicMock module for any absent third-party package above."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        mod = MagicMock(name=spec.name)
        mod.__name__ = spec.name
        mod.__path__ = []
        mod.__spec__ = spec
        mod.__loader__ = self
        return mod

    def exec_module(self, module):
        pass


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "improver"))


def load():
    """Return a namespace with the reference functions used as ground truth."""
    if not available():
        raise RuntimeError("reference tree not present")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import types

    ns = types.SimpleNamespace()
    from improver.utilities import neighbourhood_tools as nt
    from improver.utilities.rescale import rescale
    from improver.nbhood import nbhood as nb
    from improver.nbhood import radius_by_lead_time
    from improver import threshold as th

    ns.nt, ns.nb, ns.th = nt, nb, th
    ns.rescale = rescale
    ns.radius_by_lead_time = radius_by_lead_time
    return ns
