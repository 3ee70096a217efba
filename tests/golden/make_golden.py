"""Generate golden input/output vectors by running the REAL reference core.

Run in the build container only (needs /root/reference, which does not travel
to the GPU box):

    python tests/golden/make_golden.py

It imports the reference's numerical core through `_ref_loader` (iris & co.
stubbed), runs it on seeded synthetic inputs and writes
`tests/golden/nbhood_golden.npz` + `nbhood_golden.json` (case parameters).
The fixtures are committed; tests compare the oracle and the CUDA path with
them.  No reference source is copied: only its outputs are stored.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import _ref_loader  # noqa: E402

ref = _ref_loader.load()
arrays = {}
cases = []


def _store(name, arr):
    arr = np.asarray(arr)
    arrays[name] = arr


def field(kind, shape, rng):
    ny, nx = shape
    if kind == "dense":
        return rng.random(shape).astype(np.float32)
    if kind == "binary":
        return (rng.random(shape) < 0.4).astype(np.float32)
    if kind == "zeros_blob":          # mostly 0, interior blob -> e=0 trimming
        a = np.zeros(shape, np.float32)
        a[ny // 3: ny // 3 + 4, nx // 2: nx // 2 + 5] = rng.random((4, 5)).astype(np.float32)
        return a
    if kind == "ones_blob":           # mostly 1, interior blob -> e=1 trimming
        a = np.ones(shape, np.float32)
        a[ny // 2: ny // 2 + 3, nx // 3: nx // 3 + 4] = rng.random((3, 4)).astype(np.float32)
        return a
    if kind == "ones_edge":           # mostly 1, blob touching an edge -> quirk
        a = np.ones(shape, np.float32)
        a[0:2, 2:6] = rng.random((2, 4)).astype(np.float32)
        return a
    if kind == "ones_corner":
        a = np.ones(shape, np.float32)
        a[ny - 2:, nx - 3:] = 0.25
        return a
    if kind == "all_zero":
        return np.zeros(shape, np.float32)
    if kind == "all_one":
        return np.ones(shape, np.float32)
    if kind == "signed":              # not a probability: negative values
        return (rng.standard_normal(shape) * 7).astype(np.float32)
    raise KeyError(kind)


def nbhood_case(idx, method, kind, shape, cells, mask_kind, masked_input, sum_only,
                re_mask, weighted, seed):
    rng = np.random.default_rng(seed)
    data = field(kind, shape, rng)
    mask = None
    if mask_kind == "random":
        mask = (rng.random(shape) < 0.7).astype(np.int8)
    elif mask_kind == "block":        # a hole wider than the neighbourhood -> NaN outputs
        mask = np.ones(shape, np.int8)
        mask[1: 3 + 2 * cells + 2, 1: 3 + 2 * cells + 2] = 0
    inp = data
    in_mask = None
    if masked_input:
        in_mask = rng.random(shape) < 0.3
        raw = data.copy()
        raw[in_mask & (rng.random(shape) < 0.3)] = np.nan      # NaN hidden under the mask
        inp = np.ma.masked_array(raw, in_mask)
        data = raw
    plugin = ref.nb.NeighbourhoodProcessing(method, 2000.0 * cells, weighted_mode=weighted,
                                            sum_only=sum_only, re_mask=re_mask)
    if method == "circular":
        plugin.kernel = ref.nb.circular_kernel(cells, weighted)
        plugin.nb_size = max(plugin.kernel.shape)
    else:
        plugin.nb_size = 2 * cells + 1
    out = plugin._calculate_neighbourhood(inp, mask)
    name = f"nb{idx:03d}"
    _store(name + "_in", data)
    if mask is not None:
        _store(name + "_mask", mask)
    if in_mask is not None:
        _store(name + "_inmask", in_mask)
    _store(name + "_out", np.ma.getdata(out))
    if isinstance(out, np.ma.MaskedArray):
        _store(name + "_outmask", np.ma.getmaskarray(out))
    cases.append(dict(kind="nbhood", name=name, method=method, field=kind, cells=cells,
                      has_mask=mask is not None, masked_input=masked_input,
                      sum_only=sum_only, re_mask=re_mask, weighted=weighted,
                      masked_output=isinstance(out, np.ma.MaskedArray)))


class _FakeCube:
    def __init__(self, data):
        self.data = data
        self.dtype = data.dtype


def truth_case(idx, thr, bounds, ff, op, masked, seed):
    rng = np.random.default_rng(seed)
    data = (rng.gamma(2.0, 1.5, (13, 17))).astype(np.float32)
    data[0, 0] = np.float32(thr)
    data[0, 1] = np.inf
    data[0, 2] = -np.inf
    data[0, 3] = 0.0
    kw = dict(comparison_operator=op)
    if bounds is not None:
        kw["threshold_config"] = {str(thr): list(bounds)}
    else:
        kw["threshold_values"] = [thr]
        if ff is not None:
            kw["fuzzy_factor"] = ff
    plug = ref.th.Threshold(**kw)
    inp = data
    m = None
    if masked:
        m = rng.random(data.shape) < 0.25
        inp = np.ma.masked_array(data, m)
    tv = plug._calculate_truth_value(_FakeCube(inp), plug.thresholds[0], plug.fuzzy_bounds[0])
    name = f"tv{idx:03d}"
    _store(name + "_in", data)
    if m is not None:
        _store(name + "_inmask", m)
    _store(name + "_out", np.ma.getdata(tv))
    cases.append(dict(kind="truth", name=name, threshold=thr,
                      bounds=[float(b) for b in plug.fuzzy_bounds[0]], op=op, masked=masked))


def percentile_case(idx, shape, cells, pcts, kind, seed):
    rng = np.random.default_rng(seed)
    data = field(kind, shape, rng)
    kernel = ref.nb.circular_kernel(cells, False)
    # array-level restatement of the reference loop at nbhood.py:649-667 using
    # the reference's own pad_and_roll (the cube plumbing needs iris)
    win = ref.nt.pad_and_roll(data, kernel.shape, mode="mean", stat_length=max(kernel.shape) // 2)
    p = np.array(pcts, dtype=np.float32)
    out = np.empty((len(pcts),) + shape, np.float32)
    for chunk, o in zip(win, out.swapaxes(0, 1)):
        np.percentile(chunk[..., kernel > 0], p, axis=-1, out=o, overwrite_input=True)
    name = f"pc{idx:03d}"
    _store(name + "_in", data)
    _store(name + "_out", out)
    cases.append(dict(kind="percentile", name=name, cells=cells, percentiles=list(pcts), field=kind))


def main():
    idx = 0
    seed = 1000
    # square
    for kind in ["dense", "binary", "zeros_blob", "ones_blob", "ones_edge", "ones_corner",
                 "all_zero", "all_one", "signed"]:
        for cells in (1, 2, 4):
            for mask_kind in (None, "random"):
                nbhood_case(idx, "square", kind, (19, 23), cells, mask_kind, False, False, True,
                            False, seed)
                idx += 1
                seed += 1
    for kind in ["dense", "ones_edge", "zeros_blob"]:
        nbhood_case(idx, "square", kind, (19, 23), 2, "block", False, False, True, False, seed); idx += 1; seed += 1
        nbhood_case(idx, "square", kind, (19, 23), 2, None, True, False, True, False, seed); idx += 1; seed += 1
        nbhood_case(idx, "square", kind, (19, 23), 2, "random", True, False, False, False, seed); idx += 1; seed += 1
        nbhood_case(idx, "square", kind, (19, 23), 3, "random", False, True, True, False, seed); idx += 1; seed += 1
        nbhood_case(idx, "square", kind, (19, 23), 3, None, False, True, False, False, seed); idx += 1; seed += 1
    nbhood_case(idx, "square", "dense", (40, 37), 9, None, False, False, True, False, seed); idx += 1; seed += 1
    nbhood_case(idx, "square", "ones_edge", (40, 37), 9, "random", False, False, True, False, seed); idx += 1; seed += 1
    # circular
    for kind in ["dense", "binary", "zeros_blob", "ones_blob", "ones_edge", "all_one", "signed"]:
        for cells in (1, 2, 4):
            for weighted in (False, True):
                for mask_kind in (None, "random"):
                    nbhood_case(idx, "circular", kind, (19, 23), cells, mask_kind, False, False,
                                True, weighted, seed)
                    idx += 1
                    seed += 1
    for kind in ["dense", "ones_edge"]:
        nbhood_case(idx, "circular", kind, (19, 23), 2, "block", False, False, True, False, seed); idx += 1; seed += 1
        nbhood_case(idx, "circular", kind, (19, 23), 2, None, True, False, True, False, seed); idx += 1; seed += 1
        nbhood_case(idx, "circular", kind, (19, 23), 3, "random", False, True, False, True, seed); idx += 1; seed += 1
    nbhood_case(idx, "circular", "dense", (40, 37), 9, "random", False, False, True, False, seed); idx += 1; seed += 1

    # truth values
    t = 0
    for op in (">", ">=", "<", "<="):
        truth_case(t, 2.0, None, None, op, False, 2000 + t); t += 1
        truth_case(t, 2.0, None, 0.7, op, False, 2000 + t); t += 1
        truth_case(t, 2.0, (1.2, 3.5), None, op, True, 2000 + t); t += 1
    truth_case(t, -1.5, None, 0.6, ">", False, 2000 + t); t += 1
    truth_case(t, 0.0, (-0.5, 0.25), None, ">", False, 2000 + t); t += 1
    truth_case(t, 0.1, None, 0.3, ">", True, 2000 + t); t += 1

    # percentiles
    p = 0
    for kind in ("dense", "binary", "signed"):
        for cells in (1, 2, 3):
            percentile_case(p, (11, 14), cells, (5, 25, 50, 75, 95), kind, 3000 + p); p += 1
    percentile_case(p, (16, 16), 4, (0, 10, 20, 30, 40, 50, 60, 70, 80, 90, 100), "dense", 3000 + p); p += 1
    percentile_case(p, (23, 29), 10, (5, 25, 50, 75, 95), "dense", 3000 + p); p += 1

    # kernels
    for r in (1, 2, 3, 5, 10, 27):
        for w in (False, True):
            _store(f"kernel_r{r}_w{int(w)}", ref.nb.circular_kernel(r, w))
    cases.append(dict(kind="kernels", radii=[1, 2, 3, 5, 10, 27]))

    np.savez_compressed(os.path.join(HERE, "nbhood_golden.npz"), **arrays)
    with open(os.path.join(HERE, "nbhood_golden.json"), "w") as fh:
        json.dump(dict(numpy=np.__version__, cases=cases), fh, indent=0)
    print(f"{len(cases)} cases, {len(arrays)} arrays")


if __name__ == "__main__":
    main()
