"""GPU parity, scripts/patch_based_sindy.py: rows (scrambled feature view included), per-patch ridge
fit and R^2 quality, quality-weighted ensemble; goldens from the unmodified class (tests/golden/sindy.npz)."""

import numpy as np
import pytest

from oracle import sindy as OS

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    import pde_b200
    from pde_b200 import sindy

    pde_b200.load()
    return sindy


@pytest.fixture(scope="module")
def g():
    from conftest import GOLDEN

    return np.load(GOLDEN / "sindy.npz")


def test_rows_bitexact_and_scramble(S, g):
    dt, dx, dy, ps, ov = g["params"]
    U = g["images"]
    X, y = S.patch_rows(U, [(24, 48), (0, 0)], int(ps), dx, dy, dt)
    X, y = X.cpu().numpy(), y.cpu().numpy()
    for b, (y0, x0) in enumerate([(24, 48), (0, 0)]):
        seq = [f[y0:y0 + 32, x0:x0 + 32].copy() for f in U]
        Xr, yr = OS.patch_rows(seq, dx, dy, dt)
        assert np.array_equal(X[b], Xr) and np.array_equal(y[b], yr)          # reference arithmetic: bit-identical
    assert np.array_equal(X[0].reshape(6, 5, 5, 11)[0], g["lib_view_sample"])
    Xu, _ = S.patch_rows(U, [(24, 48)], int(ps), dx, dy, dt, scramble=False)
    seq = [f[24:56, 48:80].copy() for f in U]
    assert np.array_equal(Xu.cpu().numpy()[0], OS.patch_rows(seq, dx, dy, dt, scramble=False)[0])


def test_patch_fit_and_ensemble_match_the_reference(S, g):
    dt, dx, dy, ps, ov = g["params"]
    m = S.PatchBasedSINDy(dt=dt, dx=dx, dy=dy, patch_size=int(ps), overlap=int(ov))
    seq = [f[24:56, 48:80].copy() for f in g["images"]]
    c, q = m.discover_pde_for_patch(seq, alpha=0.01)
    np.testing.assert_allclose(c, g["one_coeffs"], rtol=1e-8)
    np.testing.assert_allclose(q, g["one_quality"][0], rtol=1e-8)
    m.images = [f for f in g["images"]]
    ens, names, info = m.discover_pde_patch_ensemble(alpha=0.01, min_patches=3)
    assert names == [str(n) for n in g["names"]] and info["n_patches"] == int(g["ens_n_patches"][0])
    np.testing.assert_allclose(info["patch_coeffs"], g["ens_patch_coeffs"], rtol=1e-8)
    np.testing.assert_allclose(info["patch_qualities"], g["ens_patch_qualities"], rtol=1e-8, atol=1e-12)
    assert np.array_equal(ens != 0, g["ens_coeffs"] != 0)
    np.testing.assert_allclose(ens, g["ens_coeffs"], rtol=1e-8)
    np.testing.assert_allclose(info["coeffs_std"], g["ens_std"], rtol=1e-7)
    np.testing.assert_allclose([info["avg_quality"], info["quality_std"]], g["ens_quality"], rtol=1e-8)
    # too few patches -> the reference's (None, None, {})
    assert m.discover_pde_patch_ensemble(alpha=0.01, min_patches=50) == (None, None, {})
    # a short sequence / non-finite rows
    assert m.discover_pde_for_patch(seq[:2]) == (None, 0.0)
    bad = [s.copy() for s in seq]
    bad[3][12, 12] = np.nan
    cb, qb = m.discover_pde_for_patch(bad, alpha=0.01)
    Xo, yo = OS.patch_rows(bad, dx, dy, dt)
    co, qo = OS.fit_rows(Xo, yo, 0.01)
    np.testing.assert_allclose(cb, co, rtol=1e-8)
    np.testing.assert_allclose(qb, qo, rtol=1e-8, atol=1e-12)
    with pytest.raises(NotImplementedError):
        m.discover_pde_for_patch(seq, registration_method="ecc")
