"""Loader for the committed golden vectors (tests/golden/nbhood_golden.*),
generated from the real reference by tests/golden/make_golden.py."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Golden:
    def __init__(self):
        self.arrays = np.load(os.path.join(HERE, "golden", "nbhood_golden.npz"))
        with open(os.path.join(HERE, "golden", "nbhood_golden.json")) as fh:
            self.meta = json.load(fh)
        self.cases = self.meta["cases"]

    def of_kind(self, kind):
        return [c for c in self.cases if c["kind"] == kind]

    def get(self, case, suffix):
        key = f"{case['name']}_{suffix}"
        return self.arrays[key] if key in self.arrays.files else None

    def nbhood_inputs(self, case):
        data = self.get(case, "in")
        mask = self.get(case, "mask")
        inmask = self.get(case, "inmask")
        return data, mask, inmask


_G = None


def load_golden():
    global _G
    if _G is None:
        _G = Golden()
    return _G
