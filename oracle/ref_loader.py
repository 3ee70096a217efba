"""Import the numerical core of the unmodified reference (metoppv/improver) from a directory.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (like everything under oracle/): used by
tests/golden/make_golden*.py (fixture generation from /root/reference) and by the CPU legs of
bench.py (`cpu_baseline`, `--impl reference`, from the build artefact oracle/_ref).  The product
package never imports it.

iris / cf_units / cartopy / netCDF4 are absent from the image, so they are replaced by MagicMock
modules before `improver` is imported; only array-level functions are used
afterwards (NeighbourhoodProcessing._calculate_neighbourhood, Threshold._calculate_truth_value,
pad_and_roll, circular_kernel, boxsum), which run on real numpy / scipy.
"""
import importlib.abc
import importlib.machinery
import os
import sys
import types
from unittest.mock import MagicMock

HERE = os.path.dirname(os.path.abspath(__file__))
SHIPPED_ROOT = os.path.join(HERE, "_ref")          # built by oracle/build_ref.py (git-ignored)

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


def available(root: str = SHIPPED_ROOT) -> bool:
    return os.path.isfile(os.path.join(root, "improver", "nbhood", "nbhood.py"))


def load(root: str = SHIPPED_ROOT):
    """Namespace with the reference modules used as ground truth / CPU baseline."""
    if not available(root):
        raise RuntimeError(f"no reference tree under {root}")
    if not any(isinstance(f, _StubFinder) for f in sys.meta_path):
        sys.meta_path.append(_StubFinder())
    if root not in sys.path:
        sys.path.insert(0, root)
    ns = types.SimpleNamespace()
    from improver.utilities import neighbourhood_tools as nt
    from improver.utilities.rescale import rescale
    from improver.nbhood import nbhood as nb
    from improver.nbhood import radius_by_lead_time
    from improver import threshold as th

    ns.nt, ns.nb, ns.th = nt, nb, th
    ns.rescale = rescale
    ns.radius_by_lead_time = radius_by_lead_time
    ns.root = root
    return ns
