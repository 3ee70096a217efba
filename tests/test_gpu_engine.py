"""Parity of the CUDA engine (through the C ABI) with the reference golden
vectors and with the CPU oracle.  Tolerances (BASELINE.json north_star):
max-abs 1e-5 on probabilities / means, bit-exact masks and 0/1 truth values."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import nbhood_oracle as orc  # noqa: E402
from tests.golden_cases import load_golden  # noqa: E402

G = load_golden()
ATOL = 1e-5


def _dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def _ids(cases):
    return [c["name"] for c in cases]


def _assert_close(got, exp, atol=ATOL, rtol=0.0):
    got = np.asarray(got)
    exp = np.asarray(exp)
    assert got.shape == exp.shape
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    np.testing.assert_allclose(got[ok], exp[ok], atol=atol, rtol=rtol)


@pytest.mark.parametrize("case", G.of_kind("nbhood"), ids=_ids(G.of_kind("nbhood")))
def test_neighbourhood_golden(case):
    from improver_b200 import engine

    data, mask, inmask = G.nbhood_inputs(case)
    # NaNs hidden under the input mask must not matter: keep them in the input
    out, omask, status = engine.neighbourhood(
        _dev(data), case["cells"], method=case["method"],
        mask2d=_dev(mask.astype(np.uint8)) if mask is not None else None,
        mask_nd=_dev(inmask.astype(np.uint8)) if inmask is not None else None,
        weighted_mode=case["weighted"], sum_only=case["sum_only"], want_mask=True)
    engine.check_status(status)
    exp = G.get(case, "out")
    scale = max(1.0, float(np.nanmax(np.abs(exp)))) if case["sum_only"] or case["field"] == "signed" else 1.0
    _assert_close(out.cpu().numpy(), exp, atol=ATOL * scale)
    if case["masked_output"]:
        np.testing.assert_array_equal(omask.cpu().numpy().astype(bool), G.get(case, "outmask"))


@pytest.mark.parametrize("case", G.of_kind("truth"), ids=_ids(G.of_kind("truth")))
def test_truth_value_golden_bit_exact(case):
    from improver_b200 import engine

    data = G.get(case, "in")
    m = G.get(case, "inmask")
    lo, hi = case["bounds"]
    out, contrib, status = engine.threshold(
        _dev(data), [(case["threshold"], lo, hi, case["op"])],
        mask_nd=_dev(m.astype(np.uint8)) if m is not None else None)
    engine.check_status(status)
    got = out.cpu().numpy()[0]
    exp = G.get(case, "out")
    if m is not None:
        np.testing.assert_array_equal(got[~m], exp[~m])
        np.testing.assert_array_equal(contrib.cpu().numpy(), (~m).astype(np.int32))
    else:
        np.testing.assert_array_equal(got, exp)


@pytest.mark.parametrize("case", G.of_kind("percentile"), ids=_ids(G.of_kind("percentile")))
def test_percentiles_golden(case):
    from improver_b200 import engine

    out, status = engine.neighbourhood_percentiles(_dev(G.get(case, "in")), case["cells"],
                                                   case["percentiles"])
    exp = G.get(case, "out")
    scale = max(1.0, float(np.abs(exp).max()))
    _assert_close(out.cpu().numpy(), exp, atol=ATOL * scale)


@pytest.mark.parametrize("shape,cells", [((200, 200), 5), ((97, 131), 7), ((64, 258), 27),
                                          ((33, 1042), 5), ((300, 40), 3)])
@pytest.mark.parametrize("method", ["square", "circular"])
def test_neighbourhood_vs_oracle_random(shape, cells, method):
    from improver_b200 import engine

    if method == "circular" and cells > 12:
        pytest.skip("oracle correlate too slow")
    rng = np.random.default_rng(hash((shape, cells)) % 2**32)
    data = rng.random((3,) + shape).astype(np.float32)
    data[1][data[1] < 0.6] = 0.0          # sparse slice
    data[2] = 1.0
    data[2, :4, 5:9] = 0.3                # ones with an edge blob (trimming quirk)
    mask = (rng.random(shape) < 0.8).astype(np.uint8)
    for m in (None, mask):
        out, omask, status = engine.neighbourhood(_dev(data), cells, method=method,
                                                  mask2d=_dev(m) if m is not None else None,
                                                  want_mask=True)
        engine.check_status(status)
        exp, emask = orc.neighbourhood_cube(data, m, method=method, cells=cells)
        _assert_close(out.cpu().numpy(), exp)
        np.testing.assert_array_equal(omask.cpu().numpy().astype(bool), emask)


@pytest.mark.parametrize("method", ["square", "circular"])
@pytest.mark.parametrize("fuzzy", [True, False])
def test_fused_threshold_nbhood_mean_vs_oracle(method, fuzzy):
    from improver_b200 import engine

    rng = np.random.default_rng(7)
    R, L, ny, nx = 5, 3, 61, 86
    data = rng.gamma(2.0, 1.5, (R, L, ny, nx)).astype(np.float32)
    thr = [1.0, 4.0]
    bounds = orc.fuzzy_bounds_from_factor(thr, 0.7) if fuzzy else [(t, t) for t in thr]
    cells = 4
    mask = (rng.random((ny, nx)) < 0.85).astype(np.uint8)
    for m in (None, mask):
        exp = orc.threshold_square_mean(data, thr, bounds, ">", cells, mask=m, method=method)
        out, _, status = engine.neighbourhood(
            _dev(data), cells, method=method, mask2d=_dev(m) if m is not None else None,
            thresholds=[(t, b[0], b[1], ">") for t, b in zip(thr, bounds)], collapse_members=True)
        engine.check_status(status)
        assert tuple(out.shape) == (L, len(thr), ny, nx)
        _assert_close(out.cpu().numpy(), exp)


def test_fused_quirk_slices_are_fixed_up():
    """Members that are all ones apart from an edge blob trigger the
    reference's ones-trimming edge behaviour."""
    from improver_b200 import engine

    rng = np.random.default_rng(11)
    R, L, ny, nx = 4, 2, 40, 52
    data = np.full((R, L, ny, nx), 10.0, np.float32)       # far above the threshold -> tv == 1
    data[1, 0, :3, 10:20] = rng.random((3, 10)).astype(np.float32) * 2
    data[2, 1, -2:, :5] = 0.0
    data[3, 1] = rng.gamma(2.0, 1.5, (ny, nx)).astype(np.float32)
    thr, bounds, cells = [1.0], [(0.7, 1.3)], 3
    exp = orc.threshold_square_mean(data, thr, bounds, ">", cells)
    out, _, status = engine.neighbourhood(_dev(data), cells, thresholds=[(1.0, 0.7, 1.3, ">")],
                                          collapse_members=True)
    fixed = engine.check_status(status)
    assert fixed >= 2
    _assert_close(out.cpu().numpy(), exp)


def test_nan_raises_value_error():
    from improver_b200 import engine

    data = np.random.default_rng(0).random((2, 30, 30)).astype(np.float32)
    data[1, 4, 5] = np.nan
    for method in ("square", "circular"):
        _, _, status = engine.neighbourhood(_dev(data), 2, method=method)
        with pytest.raises(ValueError, match="NaN detected"):
            engine.check_status(status)


def test_threshold_collapse_with_masks_vs_oracle():
    from improver_b200 import engine

    rng = np.random.default_rng(3)
    data = rng.gamma(2.0, 1.5, (6, 37, 41)).astype(np.float32)
    m = rng.random(data.shape) < 0.3
    m[:, 5, 5] = True                      # a point masked in every member
    thr = [0.5, 2.0, 5.0]
    bounds = orc.fuzzy_bounds_from_factor(thr, 0.8)
    exp, inv, contrib = orc.threshold_cube(np.ma.masked_array(data, m), thr, bounds, ">",
                                           collapse_leading=True)
    out, cnt, status = engine.threshold(_dev(data), [(t, b[0], b[1], ">") for t, b in zip(thr, bounds)],
                                        mask_nd=_dev(m.astype(np.uint8)), collapse_leading=True)
    engine.check_status(status)
    np.testing.assert_array_equal(cnt.cpu().numpy(), contrib.astype(np.int32))
    np.testing.assert_array_equal(out.cpu().numpy(), exp)      # same float32 accumulation order
    assert inv is not None and inv[5, 5]


@pytest.mark.parametrize("masked_input", [False, True])
def test_circular_large_radius_global_prefix_path(masked_input):
    """Radii beyond the shared-memory ring (config 2: 162 km = 81 cells) use the
    global row-prefix gather; checked here at r=45 on a small grid."""
    from improver_b200 import engine

    rng = np.random.default_rng(21)
    shape, cells = (97, 120), 45
    data = rng.random((2,) + shape).astype(np.float32)
    data[1, 10:30, 40:60] = 0.0
    mask = (rng.random(shape) < 0.7).astype(np.uint8)
    inmask = rng.random(data.shape) < 0.2 if masked_input else None
    for m in (None, mask):
        for sum_only in (False, True):
            out, omask, status = engine.neighbourhood(
                _dev(data), cells, method="circular", mask2d=_dev(m) if m is not None else None,
                mask_nd=_dev(inmask.astype(np.uint8)) if inmask is not None else None,
                sum_only=sum_only, want_mask=True)
            engine.check_status(status)
            exp = np.stack([np.ma.getdata(orc.neighbourhood(
                np.ma.masked_array(data[i], inmask[i]) if inmask is not None else data[i], m,
                method="circular", cells=cells, sum_only=sum_only, re_mask=False)) for i in range(2)])
            scale = max(1.0, float(np.nanmax(np.abs(exp)))) if sum_only else 1.0
            _assert_close(out.cpu().numpy(), exp, atol=ATOL * scale)


def test_fused_circular_large_radius():
    from improver_b200 import engine

    rng = np.random.default_rng(22)
    data = rng.gamma(2.0, 1.5, (3, 2, 90, 100)).astype(np.float32)
    exp = orc.threshold_square_mean(data, [2.0], [(1.4, 2.6)], ">", 42, method="circular")
    out, _, status = engine.neighbourhood(_dev(data), 42, method="circular",
                                          thresholds=[(2.0, 1.4, 2.6, ">")], collapse_members=True)
    engine.check_status(status)
    _assert_close(out.cpu().numpy(), exp)


def _square_oracle_stack(data, mask, cells, wrap=False, sum_only=False):
    outs = []
    for sl in data.reshape((-1,) + data.shape[-2:]):
        if wrap:
            padded = np.pad(sl, ((0, 0), (cells, cells)), mode="wrap")
            pm = np.pad(mask, ((0, 0), (cells, cells)), mode="wrap") if mask is not None else None
            res = orc.neighbourhood(padded, pm, method="square", cells=cells, sum_only=sum_only, re_mask=False)
            outs.append(np.ma.getdata(res)[:, cells:-cells])
        else:
            res = orc.neighbourhood(sl, mask, method="square", cells=cells, sum_only=sum_only, re_mask=False)
            outs.append(np.ma.getdata(res))
    return np.stack(outs).reshape(data.shape)


@pytest.mark.parametrize("shape,cells,wrap", [((5, 97, 130), 5, False), ((3, 40, 1042), 5, False),
                                              ((2, 3, 33, 58), 2, False), ((7, 64, 64), 1, False),
                                              ((2, 30, 1400), 3, False), ((3, 50, 256), 5, True),
                                              ((2, 24, 1480), 4, True), ((4, 51, 700), 9, False)])
def test_single_member_kernels_match_oracle(shape, cells, wrap):
    """Single-member slices (the un-fused square path): box kernel (radius <= 8, even widths), warp
    kernel (periodic x, radius 9) against the oracle with / without an external mask, sum_only,
    wide rows and ones-trimming slices."""
    from improver_b200 import engine

    rng = np.random.default_rng(43)
    data = rng.random(shape).astype(np.float32)
    flat = data.reshape((-1,) + shape[-2:])
    if not wrap:      # ones-trimming near the x edges has no defined meaning on a periodic axis
        flat[0] = 1.0                                        # all ones apart from an edge blob
        flat[0, :2, 4:9] = 0.25
        flat[-1] = 1.0
        flat[-1, 10:12, -3:] = 0.0
    mask = (rng.random(shape[-2:]) < 0.8).astype(np.uint8)
    for m in (None, mask):
        md = _dev(m) if m is not None else None
        for sum_only in (False, True):
            scale = max(1.0, float((2 * cells + 1) ** 2)) if sum_only else 1.0
            exp = _square_oracle_stack(data, m, cells, wrap=wrap, sum_only=sum_only)
            out, _, status = engine.neighbourhood(_dev(data), cells, mask2d=md, sum_only=sum_only, wrap_x=wrap)
            engine.check_status(status)
            _assert_close(out.cpu().numpy(), exp, atol=ATOL * scale)
    # thresholded single slices against the oracle chain with one member
    if not wrap:
        gam = rng.gamma(2.0, 1.5, (1,) + flat.shape).astype(np.float32)
        exp = orc.threshold_square_mean(gam, [2.0], [(1.4, 2.6)], ">", cells)
        out, _, status = engine.neighbourhood(_dev(gam[0]), cells, thresholds=[(2.0, 1.4, 2.6, ">")])
        engine.check_status(status)
        _assert_close(out.cpu().numpy()[:, 0], exp[:, 0])


def test_member_padded_device_layout_is_equivalent():
    """A fused cube whose member blocks are padded apart (pipeline.member_padded_empty, the layout
    bench.py and the host pipeline use) gives bit-identical results to the contiguous cube."""
    from improver_b200 import engine, pipeline

    rng = np.random.default_rng(51)
    data = rng.gamma(2.0, 1.5, (6, 3, 70, 130)).astype(np.float32)
    cube = pipeline.member_padded_empty(data.shape, "cuda", pad_bytes=4096)
    assert not cube.is_contiguous() and cube[0].is_contiguous()
    cube.copy_(torch.from_numpy(data))
    thr = [(2.0, 1.4, 2.6, ">")]
    a, _, st = engine.neighbourhood(cube, 4, thresholds=thr, collapse_members=True)
    engine.check_status(st)
    b, _, st = engine.neighbourhood(_dev(data), 4, thresholds=thr, collapse_members=True)
    engine.check_status(st)
    np.testing.assert_array_equal(a.cpu().numpy(), b.cpu().numpy())
    _assert_close(a.cpu().numpy(), orc.threshold_square_mean(data, [2.0], [(1.4, 2.6)], ">", 4))


def test_graphed_call_replays_the_fused_chain():
    """pipeline.GraphedCall (what bench.py times): the fused call captured in a CUDA graph gives
    the same result as the eager call, also after the input buffer is refilled."""
    from improver_b200 import engine, pipeline

    rng = np.random.default_rng(71)
    data = rng.gamma(2.0, 1.5, (6, 2, 80, 130)).astype(np.float32)
    cube = _dev(data)
    out = torch.empty((2, 1, 80, 130), dtype=torch.float32, device="cuda")
    thr = [(2.0, 1.4, 2.6, ">")]

    def step():
        return engine.neighbourhood(cube, 4, thresholds=thr, collapse_members=True, out=out)[2]

    graphed = pipeline.GraphedCall(step)
    engine.check_status(graphed())
    _assert_close(out.cpu().numpy(), orc.threshold_square_mean(data, [2.0], [(1.4, 2.6)], ">", 4))
    data2 = rng.gamma(2.0, 1.5, data.shape).astype(np.float32)
    cube.copy_(torch.from_numpy(data2))
    engine.check_status(graphed())
    _assert_close(out.cpu().numpy(), orc.threshold_square_mean(data2, [2.0], [(1.4, 2.6)], ">", 4))
