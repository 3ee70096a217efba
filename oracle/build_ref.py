#!/usr/bin/env python
"""Recipe for oracle/_ref: a verbatim copy of the reference's `improver/` Python package
(sources only) taken from /root/reference in the build container, so that the CPU legs of
bench.py can time the REFERENCE'S OWN functions on the GPU box, where /root/reference does not
exist.  oracle/_ref is a build artefact: git-ignored (never committed), not gpurun-ignored (it
travels with the snapshot like the built .so).  Run by __graft_entry__.build() when the
reference tree is present; `python oracle/build_ref.py` does the same by hand."""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.environ.get("IMPROVER_REFERENCE", "/root/reference")
DST = os.path.join(HERE, "_ref")


def build(force: bool = False) -> str:
    src_pkg = os.path.join(SRC, "improver")
    if not os.path.isdir(src_pkg):
        raise RuntimeError(f"{src_pkg} not found: oracle/_ref can only be built where the reference is")
    manifest = os.path.join(DST, "MANIFEST.json")
    if os.path.exists(manifest) and not force:
        return DST
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    digest = hashlib.sha256()
    n = 0
    for root, dirs, files in os.walk(src_pkg):
        dirs[:] = sorted(d for d in dirs if d != "__pycache__")
        for f in sorted(files):
            if not f.endswith(".py"):
                continue
            s = os.path.join(root, f)
            rel = os.path.relpath(s, SRC)
            d = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copyfile(s, d)
            with open(s, "rb") as fh:
                digest.update(rel.encode() + b"\0" + fh.read())
            n += 1
    with open(manifest, "w") as fh:
        json.dump({"source": SRC, "files": n, "sha256": digest.hexdigest(),
                   "note": "unmodified copy of the reference's improver/ package (python sources)"}, fh, indent=1)
    return DST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
