"""GPU parity, analyze_results dialect: slice-aligned differences with the central time difference,
the six nested model libraries from ONE K1 pass, the script's scikit-learn STRidge, and its validation helpers
(one-step check, k-step rollouts).  Goldens: tests/golden/analyze.npz, produced by executing the reference's own
source lines (make_golden.golden_analyze); bars: statistics 1e-10, identical support, coefficients 1e-8."""

import numpy as np
import pytest

from helpers import assert_coef_close, assert_stats_close
from oracle import analyze as OA
from oracle import gram

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def A():
    import pde_b200
    from pde_b200 import analyze

    pde_b200.load()
    return analyze


@pytest.fixture(scope="module")
def g():
    from conftest import GOLDEN

    return np.load(GOLDEN / "analyze.npz")


def test_statistics_of_the_13_term_library(A, g):
    from pde_b200 import _lib as L
    from pde_b200 import ops

    U = g["U"]
    dx, dy, dt = g["spacing"]
    stats, tr, te = A.model_stats(U, dx, dy, dt, 0.7)
    assert tr.stop == int(g["train_stop"][0]) and te.stop == int(g["aligned"][0])
    d = OA.derivatives(U, dx, dy, dt)
    for f, sl in enumerate((tr, te)):
        X, y = OA.rows(d, OA.FULL_NAMES, sl)
        assert_stats_close(stats[f].cpu().numpy(), gram.pack_stats(X, y), 13)
    # the sampled-rows entry point in this dialect: rows bit-identical to the reference's slices (u^3 to 1 ulp)
    idx = np.random.default_rng(0).choice(d["u"].size, size=300, replace=False)
    X, y = ops.fd_gather_rows(U, dy, dx, dt, idx, dialect=L.FD_SLICE_CENTRAL, library=L.LIB_AR_FULL)
    Xr, yr = OA.rows(d, OA.FULL_NAMES, slice(0, None))
    X, y = X.cpu().numpy(), y.cpu().numpy()
    assert np.array_equal(y, yr[idx])
    for k, n in enumerate(OA.FULL_NAMES):
        if n == "u^3":
            np.testing.assert_allclose(X[:, k], Xr[idx, k], rtol=3e-16)
        else:
            assert np.array_equal(X[:, k], Xr[idx, k]), n
    # block means exist in this dialect too (generic kernel)
    st = ops.fd_lib_gram(U, dy, dx, dt, dialect=L.FD_SLICE_CENTRAL, library=L.LIB_AR_FULL, block=(2, 4, 5)).cpu().numpy()[0]
    Xf, yf = OA.rows(d, OA.FULL_NAMES, slice(0, None))
    T, H, W = d["u"].shape
    Xb = Xf.reshape(T, H, W, 13)[:, :, :, :]
    rows_x, rows_y = [], []
    for t0 in range(0, T, 2):
        for i0 in range(0, H, 4):
            for j0 in range(0, W, 5):
                rows_x.append(Xb[t0:t0 + 2, i0:i0 + 4, j0:j0 + 5].reshape(-1, 13).mean(axis=0))
                rows_y.append(yf.reshape(T, H, W)[t0:t0 + 2, i0:i0 + 4, j0:j0 + 5].mean())
    assert_stats_close(st, gram.pack_stats(np.array(rows_x), np.array(rows_y)), 13)


def test_models_1_to_6_match_the_reference(A, g):
    dx, dy, dt = g["spacing"]
    out = A.fit_models(g["U"], dx, dy, dt, train_frac=0.7, alpha=0.01, threshold=1e-5)
    assert list(out) == list(OA.MODELS)
    for idx, (name, r) in enumerate(out.items(), start=1):
        assert r["names"] == [str(n) for n in g[f"m{idx}_names"]]
        assert_coef_close(r["coeffs"], g[f"m{idx}_coeffs"], what=name)                 # identical support, 1e-8
        np.testing.assert_allclose(r["scale"], g[f"m{idx}_scale"], rtol=1e-10)
        np.testing.assert_allclose([r["train"]["r2"], r["train"]["rmse"]], g[f"m{idx}_train_metrics"][:2], rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose([r["test"]["r2"], r["test"]["rmse"]], g[f"m{idx}_test_metrics"][:2], rtol=1e-8, atol=1e-10)


def test_stridge_signature_and_grid(A, g):
    dx, dy, dt = g["spacing"]
    d = OA.derivatives(g["U"], dx, dy, dt)
    tr = slice(0, int(g["train_stop"][0]))
    X, y = OA.rows(d, OA.FULL_NAMES, tr)
    assert np.array_equal(X[:5], g["m6_X_train_head"])
    for (a, th), ref in zip(g["m6_grid"], g["m6_grid_coeffs"]):
        c, scaler = A.stridge(X, y, alpha=a, threshold=th)
        assert_coef_close(c, ref, what=f"analyze stridge {a} {th}")
    np.testing.assert_allclose(scaler.scale_, g["m6_scale"], rtol=1e-10)
    X3, y3 = OA.rows(d, OA.MODELS["Model 3: + First order spatial"], tr)
    c3, _ = A.stridge(X3, y3)
    assert_coef_close(c3, g["m3_coeffs"], what="model 3 through the literal stridge")
    assert A.split_time(14, 0.7) == OA.split_time(14, 0.7)
    with pytest.raises(ValueError):
        A.split_time(14, 0.95)


def test_one_step_and_rollout_checks(A, g):
    dx, dy, dt = g["spacing"]
    d = OA.derivatives(g["U"], dx, dy, dt)
    tr, te = OA.split_time(d["u"].shape[0], 0.7)
    u = np.ascontiguousarray(d["u"])
    for idx in (3, 6):
        names = [str(n) for n in g[f"m{idx}_names"]]
        c = g[f"m{idx}_coeffs"]
        Xtr, _ = OA.rows(d, names, tr)
        Xte, _ = OA.rows(d, names, te)
        ut = np.zeros_like(d["u_t"])
        ut[tr] = (Xtr @ c).reshape(d["u_t"][tr].shape)
        ut[te] = (Xte @ c).reshape(d["u_t"][te].shape)
        got = [A.one_step_prediction_rmse(u[tr], ut[tr], dt=dt), A.one_step_prediction_rmse(u[te], ut[te], dt=dt)]
        np.testing.assert_allclose(got, g[f"m{idx}_one_step"], rtol=1e-12)
        ro = np.array([[A.rollout_k_rmse(u, names, c, k, sl, dx=dx, dy=dy, dt=dt)[q] for q in ("rmse", "nrmse")]
                       for k in (1, 3) for sl in (tr, te)])
        np.testing.assert_allclose(ro, g[f"m{idx}_rollout"], rtol=1e-11)
    # spatial mask, degenerate slices, unsupported term
    mask = np.zeros(u.shape[1:], dtype=bool)
    mask[:, :12] = True
    names = [str(n) for n in g["m3_names"]]
    got = A.rollout_k_rmse(u, names, g["m3_coeffs"], 2, te, mask, dx=dx, dy=dy, dt=dt)
    ref = OA.rollout_k_rmse(u, names, g["m3_coeffs"], 2, te, dx, dy, dt, mask)
    np.testing.assert_allclose([got["rmse"], got["nrmse"]], [ref["rmse"], ref["nrmse"]], rtol=1e-11)
    assert np.isnan(A.rollout_k_rmse(u, names, g["m3_coeffs"], 0, te, dx=dx, dy=dy, dt=dt)["rmse"])
    assert np.isnan(A.rollout_k_rmse(u, names, g["m3_coeffs"], 9, te, dx=dx, dy=dy, dt=dt)["rmse"])
    with pytest.raises(KeyError):
        A.rollout_k_rmse(u, ["u", "u_xy"], [1.0, 1.0], 1, tr, dx=dx, dy=dy, dt=dt)
    m1 = A.one_step_prediction_rmse(u[tr], np.zeros_like(u[tr]), dt=dt, spatial_mask=mask)
    np.testing.assert_allclose(m1, OA.one_step_prediction_rmse(u[tr], np.zeros_like(u[tr]), dt, mask), rtol=1e-12)


def test_spatial_holdout_and_larger_stack(A):
    """Left / right spatial hold-out (ar:282-299) through per-row folds, on a stack large enough to fill the GPU."""
    from pde_b200 import ops

    U = ops.synth_field(12, 70, 90, seed=4, kind=1, noise=0.02).cpu().numpy()
    mask = np.zeros((68, 88), dtype=bool)
    mask[:, :int(np.floor(0.7 * 88))] = True
    out = A.fit_models(U, 0.1, 0.1, 1.0, spatial_mask=mask)
    d = OA.derivatives(U, 0.1, 0.1, 1.0)
    m3 = m3b = np.broadcast_to(mask, d["u"].shape)
    for name, names in list(OA.MODELS.items())[2:4]:
        terms = OA.library_terms(d)
        Xtr = np.column_stack([terms[n][m3].ravel() for n in names])
        ytr = d["u_t"][m3b].ravel()
        c, _ = OA.stridge(Xtr, ytr)
        assert_coef_close(out[name]["coeffs"], c, what=name)
