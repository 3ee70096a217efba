"""numpy restatements of the arithmetic the fixed-point kernels rely on (CPU): the conversion
trick of the box kernel, the correctly rounded quotient, the fixed-point window sums against the
oracle's float64 box sums, and the periodic-x prefix formula of the H pass.  These pin the
ALGORITHMS; the kernels themselves are checked against the oracle in the -m gpu tests."""
import numpy as np
import pytest

from oracle import nbhood_oracle as orc

F32 = np.float32


def _fma32(a, b, c):
    """float32 fused multiply-add of float32 arrays (exact product and sum in float64 here: the
    operands are a float32 and a power of two, or 24-bit integers)."""
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F32)


@pytest.mark.parametrize("shift", [12, 17, 20, 22])
def test_magic_number_conversion_equals_round_half_even(shift):
    """improver_b200/csrc/nbh_square_box.cuh box_to_fixed: tv * 2^s + 1.5 * 2^23 leaves
    rint(tv * 2^s) in the low mantissa bits for 0 <= tv * 2^s <= 2^22."""
    rng = np.random.default_rng(shift)
    tv = np.concatenate([rng.random(1_000_000).astype(F32),
                         np.array([0, 1, 0.5, 2.0 ** -23, 3 * 2.0 ** -23, 1.5 * 2.0 ** -shift, 2.5 * 2.0 ** -shift,
                                   1 - 2.0 ** -24], dtype=F32)])
    scale = np.full_like(tv, 2.0 ** shift)
    magic = _fma32(tv, scale, np.full_like(tv, 12582912.0))
    q = magic.view(np.uint32) - np.uint32(0x4B400000)
    ref = np.rint(tv.astype(np.float64) * 2.0 ** shift).astype(np.uint32)
    np.testing.assert_array_equal(q, ref)


def test_reciprocal_plus_two_fmas_is_the_correctly_rounded_quotient():
    """S / n with an approximate reciprocal (here: the float32 reciprocal perturbed by up to 2 ulp,
    as rcp.approx may be) followed by q0 = S * rc, rem = fma(-q0, n, S), q = fma(rem, rc, q0)."""
    rng = np.random.default_rng(3)
    n_cells = rng.integers(1, 121 * 18 + 1, 200_000)
    shift = 20
    n = (n_cells.astype(np.float64) * 2.0 ** shift).astype(F32)              # exact: small integer times 2^s
    S = (rng.integers(0, n_cells + 1) * 2.0 ** shift * rng.random(n_cells.size)).astype(np.uint32).astype(F32)
    rc = (1.0 / n.astype(np.float64)).astype(F32)
    rc = np.nextafter(rc, np.where(rng.random(rc.size) < 0.5, F32(0), F32(1))).astype(F32)
    rc = np.nextafter(rc, np.where(rng.random(rc.size) < 0.5, F32(0), F32(1))).astype(F32)
    q0 = (S * rc).astype(F32)
    rem = _fma32(-q0, n, S)
    q = _fma32(rem, rc, q0)
    exact = (S.astype(np.float64) / n.astype(np.float64)).astype(F32)
    np.testing.assert_array_equal(q, exact)


@pytest.mark.parametrize("cells", [1, 5, 8])
def test_fixed_point_window_sums_match_the_oracle(cells):
    """q = rint(v * 2^s) with (2r+1)^2 * 2^s < 2^32 (s <= 22), unsigned 32-bit window sums that wrap,
    correctly rounded S / (n * 2^s): within 2^-(s+1) of the oracle and exact for constants."""
    rng = np.random.default_rng(cells)
    ny, nx = 40, 56
    v = rng.random((ny, nx)).astype(F32)
    nb = 2 * cells + 1
    s = 0
    while s < 22 and nb * nb * 2.0 ** (s + 1) < 2.0 ** 32:
        s += 1
    q = np.rint(v.astype(np.float64) * 2.0 ** s).astype(np.uint32)
    pad = np.zeros((ny + 2 * cells + 1, nx + 2 * cells + 1), np.uint32)
    pad[cells + 1:cells + 1 + ny, cells + 1:cells + 1 + nx] = q
    P = pad.cumsum(0, dtype=np.uint32).cumsum(1, dtype=np.uint32)           # wraps modulo 2^32
    S = (P[nb:, nb:] - P[:-nb, nb:] - P[nb:, :-nb] + P[:-nb, :-nb]).astype(np.uint32)
    yy, xx = np.arange(ny)[:, None], np.arange(nx)[None, :]
    cnt = ((np.minimum(yy + cells, ny - 1) - np.maximum(yy - cells, 0) + 1) *
           (np.minimum(xx + cells, nx - 1) - np.maximum(xx - cells, 0) + 1))
    got = (S.astype(np.float64) / (cnt * 2.0 ** s)).astype(F32)
    exp = np.ma.getdata(orc.neighbourhood(v, None, method="square", cells=cells, re_mask=False))
    assert np.max(np.abs(got.astype(np.float64) - exp)) <= 2.0 ** -(s + 1) + 2.0 ** -24
    const = np.full((ny, nx), 0.3, F32)
    qc = np.rint(const.astype(np.float64) * 2.0 ** s)
    assert np.all(((qc * cnt) / (cnt * 2.0 ** s)).astype(F32) == (qc[0, 0] / 2.0 ** s).astype(F32))


@pytest.mark.parametrize("nx,r", [(16, 3), (40, 5), (64, 27)])
def test_periodic_row_window_from_the_inclusive_prefix(nx, r):
    """H pass of the two-pass kernels with NBH_WRAP_X (improver_b200/csrc/nbh_square_vh.cu):
    S(x) = P[x+r] - P[x-r-1] inside the row, P[x+r] + P[nx-1] - P[nx+x-r-1] where the window
    leaves the row on the left, P[nx-1] + P[x+r-nx] - P[x-r-1] on the right (uint32, wrapping)."""
    rng = np.random.default_rng(nx + r)
    v = rng.integers(0, 2 ** 30, nx, dtype=np.uint64).astype(np.uint32)
    P = np.cumsum(v, dtype=np.uint32)
    total = P[nx - 1]
    got = np.empty(nx, np.uint32)
    with np.errstate(over="ignore"):                       # the wrap-around is the point
        for x in range(nx):
            lo = x - r - 1
            if x + r >= nx:
                got[x] = total + P[x + r - nx] - P[lo]
            elif lo < 0:
                got[x] = P[x + r] + total - P[nx + lo]
            else:
                got[x] = P[x + r] - P[lo]
    exp = np.array([np.take(v, np.arange(x - r, x + r + 1), mode="wrap").sum(dtype=np.uint32) for x in range(nx)],
                   dtype=np.uint32)
    if nx >= 2 * r + 2:
        np.testing.assert_array_equal(got, exp)
    else:
        pytest.skip("a window wraps at one row end at most (the planner rejects narrower rows)")
