"""The algebraic identities the CUDA kernels rely on, checked on the CPU against the oracle's literal restatement of
the reference (no GPU): what DESIGN.md calls the statistics formulation, the two-stage block route, the telescoped
time derivative and the zero sums of periodic difference stencils."""

import numpy as np
import pytest

from helpers import assert_stats_close, ks_rows
from oracle import gram
from oracle import ks2d as O


def _field(shape, seed):
    rng = np.random.default_rng(seed)
    T, A0, A1 = shape
    t, i, j = np.meshgrid(np.arange(T), np.arange(A0), np.arange(A1), indexing="ij")
    u = (0.5 * np.sin(2 * np.pi * (3 * i / A0 + 2 * j / A1) - 0.3 * t) + 0.3 * np.cos(2 * np.pi * (5 * i / A0 - j / A1) + 0.1 * t))
    return u + 0.05 * rng.standard_normal(shape)


@pytest.mark.parametrize("seed", range(4))
def test_statistics_are_additive_over_row_sets(seed):
    """Time slabs, tiles, folds and ranks all rely on this: statistics of a union of rows = sum of statistics."""
    rng = np.random.default_rng(seed)
    X, y = rng.standard_normal((500, 5)), rng.standard_normal(500)
    cut = int(rng.integers(1, 499))
    whole = gram.pack_stats(X, y)
    parts = gram.pack_stats(X[:cut], y[:cut]) + gram.pack_stats(X[cut:], y[cut:])
    np.testing.assert_allclose(parts, whole, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("shape,block", [((10, 32, 48), (3, 16, 16)), ((9, 40, 56), (2, 16, 24)), ((7, 24, 72), (3, 24, 32))])
def test_block_means_of_whole_sub_blocks(shape, block):
    """Two-stage route (api.cu: two_stage_blocks): the row of a (bt, 8m, 8n) block is the mean of the rows of its (bt, 8, 8)
    sub-blocks, ragged edge blocks included (ks2d:380-389 keeps them with their own divisor); nonlinear columns too."""
    U = _field(shape, 1)
    names, X8, y8 = ks_rows(U, 0.5, 0.4, 1e-2, "rich", False, (block[0], 8, 8))
    _, Xb, yb = ks_rows(U, 0.5, 0.4, 1e-2, "rich", False, block)
    T, A0, A1 = shape
    nbt, s0, s1 = -(-(T - 1) // block[0]), A0 // 8, A1 // 8
    rows8 = np.concatenate([y8[:, None], X8], axis=1).reshape(nbt, s0, s1, -1)
    m, n = block[1] // 8, block[2] // 8
    out = []
    for tb in range(nbt):
        for ib in range(-(-s0 // m)):
            for jb in range(-(-s1 // n)):
                sub = rows8[tb, ib * m:(ib + 1) * m, jb * n:(jb + 1) * n].reshape(-1, rows8.shape[-1])
                out.append(sub.mean(axis=0))
    out = np.array(out)
    np.testing.assert_allclose(out[:, 0], yb, rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(out[:, 1:], Xb, rtol=1e-11, atol=1e-13)


def test_time_derivative_of_a_block_needs_only_block_sums_of_the_next_frame():
    """pg_fd_lib_gram_tail: replacing the trailing frame by any frame with the same (8, 8) block means leaves every row
    unchanged (the block mean of the forward difference telescopes; the other columns never read that frame)."""
    U = _field((8, 16, 32), 2)
    _, X, y = ks_rows(U, 0.5, 0.5, 1e-2, "true", False, (3, 8, 8))
    V = U.copy()
    means = U[-1].reshape(2, 8, 4, 8).mean(axis=(1, 3))
    V[-1] = np.kron(means, np.ones((8, 8)))
    _, X2, y2 = ks_rows(V, 0.5, 0.5, 1e-2, "true", False, (3, 8, 8))
    assert np.array_equal(X, X2)
    np.testing.assert_allclose(y2, y, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("dictionary,adv", [("true", True), ("rich", False)])
def test_periodic_difference_stencil_columns_sum_to_rounding_noise(dictionary, adv):
    """Pointwise KS kernel: the sum over whole periodic frames of lap, bih, u_x, u_y is zero up to rounding in the
    reference's own arithmetic, so the kernel emits exact zeros for these linear sums."""
    U = _field((6, 24, 40), 3)
    names, X, y = ks_rows(U, 0.5, 0.4, 1e-2, dictionary, adv, (1, 1, 1))
    stencil = [k for k, nm in enumerate(names) if nm in ("u_x", "u_y", "∇²u", "∇⁴u")]
    assert len(stencil) == 4
    for k in stencil:
        col = X[:, k]
        assert abs(col.sum()) <= 1e-10 * np.sqrt(len(col) * (col ** 2).sum())
    p = len(names)
    s = gram.pack_stats(X, y)
    z = s.copy()
    for k in stencil:
        z[3 + k] = 0.0                      # what the kernel emits
    assert_stats_close(z, s, p)


def test_linear_sum_of_a_product_column_is_a_pair_sum():
    """Pointwise kernel: sum(u^2 * 1) == sum(u * u) and sum((u lap) * 1) == sum(u * lap): read from the pair accumulators."""
    U = _field((5, 16, 24), 4)
    names, X, y = ks_rows(U, 0.5, 0.4, 1e-2, "rich", False, (1, 1, 1))
    iu, iu2, il, iul = names.index("u"), names.index("u^2"), names.index("∇²u"), names.index("u·∇²u")
    np.testing.assert_allclose(X[:, iu2].sum(), (X[:, iu] * X[:, iu]).sum(), rtol=1e-12)
    np.testing.assert_allclose(X[:, iul].sum(), (X[:, iu] * X[:, il]).sum(), rtol=1e-12)


def test_block_sum_of_u_lap_from_neighbour_pair_products():
    """Blockwise rich library (tiled.cu, march_frame): sum over a block of u * L' with
    L' = rho (u[i+1] + u[i-1]) + (u[j+1] + u[j-1]) + kappa u  equals  kappa sum u^2 + rho * (vertical neighbour pairs)
    + (horizontal neighbour pairs), boundary pairs counted once and interior pairs twice."""
    rng = np.random.default_rng(5)
    w = rng.standard_normal((12, 8))                      # band rows 0..11, window columns own-2 .. own+5
    rho = 1.7
    kappa = -2.0 * (1.0 + rho)
    own = w[2:10, 2:6]
    Lp = rho * (w[3:11, 2:6] + w[1:9, 2:6]) + (w[2:10, 3:7] + w[2:10, 1:5]) + kappa * own
    direct = (own * Lp).sum()
    vb = (w[1, 2:6] * w[2, 2:6]).sum() + (w[9, 2:6] * w[10, 2:6]).sum()
    vi = sum((w[s, 2:6] * w[s + 1, 2:6]).sum() for s in range(2, 9))
    he = (w[2:10, 1] * w[2:10, 2]).sum() + (w[2:10, 5] * w[2:10, 6]).sum()
    hi = sum((w[2:10, q] * w[2:10, q + 1]).sum() for q in (2, 3, 4))
    pairs = kappa * (own ** 2).sum() + rho * (vb + 2 * vi) + (he + 2 * hi)
    np.testing.assert_allclose(pairs, direct, rtol=1e-12)


def test_block_sums_of_linear_terms_from_boundary_scalars():
    """Blockwise kernel (tiled.cu, march_frame): by the discrete divergence theorem the block sums of L' and of
    L'(L') (the unscaled lap and bih) reduce to per-row scalars rsU, D, e of the lane's 12 x 8 window -- the formulas
    of the kernel, transcribed, against the direct sums."""
    rng = np.random.default_rng(6)
    w = rng.standard_normal((12, 8))
    rho = 0.64
    kappa = -2.0 * (1.0 + rho)

    def Lp(a):      # L' on the interior of a 2-D array
        return rho * (a[2:, 1:-1] + a[:-2, 1:-1]) + (a[1:-1, 2:] + a[1:-1, :-2]) + kappa * a[1:-1, 1:-1]

    L1 = Lp(w)                       # rows 1..10, columns 1..6
    direct_SL = L1[1:9, 1:5].sum()   # own rows 2..9, own columns 2..5
    direct_bih = Lp(L1)[0:8, 0:4].sum()   # L'(L') on rows 2..9, columns 2..5

    rsU = w[:, 2:6].sum(axis=1)
    D = (w[:, 1] - w[:, 2]) + (w[:, 6] - w[:, 5])
    e = (w[:, 0] - w[:, 3]) + (w[:, 7] - w[:, 4])
    sD, sE = D[2:10].sum(), e[2:10].sum()
    rsL = lambda s: rho * ((rsU[s + 1] + rsU[s - 1]) - 2.0 * rsU[s]) + D[s]   # noqa: E731
    SL = rho * ((rsU[10] - rsU[9]) - (rsU[2] - rsU[1])) + sD
    SE1 = (rsL(1) - rsL(2)) + (rsL(10) - rsL(9))
    SE2 = rho * ((D[10] + D[1]) - (D[2] + D[9])) + (-3.0 * sD + sE)
    np.testing.assert_allclose(SL, direct_SL, rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(rho * SE1 + SE2, direct_bih, rtol=1e-11, atol=1e-11)
