"""Import the read-only reference's numerical core in THIS container (fixture generation and
the optional `-m "not gpu"` cross-check tests, which skip when /root/reference is absent: it does
not exist on the GPU box).  The loader itself lives in oracle/ref_loader.py."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402

REF_ROOT = os.environ.get("IMPROVER_REFERENCE", "/root/reference")


def available() -> bool:
    return ref_loader.available(REF_ROOT)


def load():
    """Return a namespace with the reference functions used as ground truth."""
    if not available():
        raise RuntimeError("reference tree not present")
    return ref_loader.load(REF_ROOT)
