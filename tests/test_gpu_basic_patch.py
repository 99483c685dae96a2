"""GPU parity, basic_usage and patch dialects (through the C ABI), against the golden
fixtures generated from the unmodified reference and the oracle on the same inputs."""

import numpy as np
import pytest

from helpers import COEF_RTOL, assert_coef_close, assert_stats_close, synthetic_stack
from oracle import basic as OB
from oracle import gram
from oracle import patch as OP

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def B():
    from pde_b200 import basic_usage

    return basic_usage


@pytest.fixture(scope="module")
def P():
    from pde_b200 import patch

    return patch


@pytest.fixture(scope="module")
def L():
    from pde_b200 import _lib

    return _lib


@pytest.fixture(scope="module")
def ops():
    from pde_b200 import ops

    return ops


# --------------------------------------------------------------------------- basic_usage
def test_basic_compute_derivatives_bitexact(B, golden_basic):
    g = golden_basic
    ut, u, ux, uy, lap = B.compute_derivatives(g["small_u"], *g["small_d"])
    assert ut.shape == g["small_ut"].shape
    for a, k in [(ut, "small_ut"), (ux, "small_ux"), (uy, "small_uy"), (lap, "small_lap")]:
        assert np.array_equal(a, g[k]), k
    assert np.array_equal(u, g["small_u"][:-1, 2:-2, 2:-2])
    Theta, names = B.build_library(u, ux, uy, lap)
    assert names == OB.TERM_NAMES and np.array_equal(Theta, g["small_Theta"])


def test_basic_stridge_regression(B, golden_basic):
    g = golden_basic
    ut = OB.compute_derivatives(g["small_u"], *g["small_d"])[0].reshape(-1)
    for (a, t), ref in zip(g["small_grid"], g["small_coef"]):
        assert_coef_close(B.stridge_regression(g["small_Theta"], ut, alpha=a, threshold=t), ref, what=f"basic {a} {t}")
    assert np.array_equal(B.stridge_regression(g["small_Theta"], ut, max_iter=0), g["small_coef_iter0"])
    with pytest.raises(ValueError):
        B.stridge_regression(g["small_Theta"], ut[:-1])


def test_basic_default_example_materialised_and_fused(B, golden_basic):
    """examples/basic_usage.py main(): 30x60x60 synthetic stack, alpha = threshold = 0.01."""
    g = golden_basic
    u, x, y, t = OB.generate_synthetic_data(n_frames=30, h=60, w=60)
    dx, dy, dt = g["default_spacing"]
    ut, uu, ux, uy, lap = B.compute_derivatives(u, dx, dy, dt)
    Theta, _ = B.build_library(uu, ux, uy, lap)
    assert Theta.shape == tuple(g["default_Theta_shape"])
    c = B.stridge_regression(Theta, ut.flatten(), alpha=0.01, threshold=0.01)
    assert_coef_close(c, g["default_coef"], what="basic default (materialised)")
    out = B.fit_from_field(u, dx, dy, dt, alpha=0.01, threshold=0.01)
    assert_coef_close(out["coef"], g["default_coef"], what="basic default (fused)")
    d = OB.compute_derivatives(u, dx, dy, dt)
    ref = gram.pack_stats(OB.build_library(*d[1:])[0], d[0].reshape(-1))
    assert_stats_close(out["stats"][0], ref, 6)


@pytest.mark.parametrize("shape", [(2, 5, 5), (3, 5, 9), (4, 17, 6), (5, 33, 40)])
def test_basic_fused_shapes_and_folds(ops, L, shape):
    rng = np.random.default_rng(shape[1])
    u = rng.standard_normal(shape)
    dx, dy, dt = 0.3, 0.7, 0.05
    ut, uu, ux, uy, lap = OB.compute_derivatives(u, dx, dy, dt)
    Theta, _ = OB.build_library(uu, ux, uy, lap)
    s = ops.fd_lib_gram(u, dy, dx, dt, dialect=L.FD_BASIC_TRIM, library=L.LIB_BASIC).cpu().numpy()
    assert_stats_close(s[0], gram.pack_stats(Theta, ut.reshape(-1)), 6)
    if shape[0] > 2:
        fof = (np.arange(shape[0] - 1) % 2).astype(np.int32)
        s = ops.fd_lib_gram(u, dy, dx, dt, dialect=L.FD_BASIC_TRIM, library=L.LIB_BASIC, fold_of_frame=fof, n_folds=2).cpu().numpy()
        per = (shape[1] - 4) * (shape[2] - 4)
        rf = np.repeat(fof, per)
        for f in range(2):
            assert_stats_close(s[f], gram.pack_stats(Theta[rf == f], ut.reshape(-1)[rf == f]), 6)


def test_basic_degenerate_and_errors(B, L):
    assert B.compute_derivatives(np.ones((3, 4, 9)), 1, 1, 1)[0].shape == (2, 0, 5)
    with pytest.raises(ValueError):
        B.compute_derivatives(np.ones((4, 4)), 1, 1, 1)
    import pde_b200

    with pytest.raises(pde_b200.PdeGramError, match="A0, A1 >= 5"):
        pde_b200.ops.fd_lib_gram(np.ones((3, 4, 9)), 1, 1, 1, dialect=L.FD_BASIC_TRIM, library=L.LIB_BASIC)


# --------------------------------------------------------------------------- patch
def test_patch_derivatives_and_rows(P, golden_patch):
    g = golden_patch
    U = g["U"]
    d = np.array([P.local_poly_derivatives(U, *p, 2, 3, 3, 1.0, 0.1, 0.1) for p in g["pts"][:3]])
    np.testing.assert_allclose(d, g["derivs"][:3], rtol=1e-8, atol=1e-11)   # atol: second derivatives that are ~0 at a point
    d2 = np.array([P.local_poly_derivatives(U, *p, 1, 2, 2, 0.5, 0.2, 0.3) for p in g["pts"][:3]])
    np.testing.assert_allclose(d2, g["derivs_deg2_r1"][:3], rtol=1e-8, atol=1e-11)
    X, y = P.build_dataset(U, [tuple(p) for p in g["pts"]], 2, 3, 3, 1.0, 0.1, 0.1, P.Library(names=P.FULL_NAMES))
    np.testing.assert_allclose(X, g["X8"], rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(y, g["y8"], rtol=1e-8, atol=1e-10)
    X6, _ = P.build_dataset(U, [tuple(p) for p in g["pts"]], 2, 3, 3, 1.0, 0.1, 0.1, P.Library(names=P.MODEL4_NAMES))
    np.testing.assert_allclose(X6, g["X6"], rtol=1e-8, atol=1e-11)
    # float64 stack gives the same rows as the float32 one up-cast (patch:220)
    X64, y64 = P.build_dataset(U.astype(np.float64), [tuple(p) for p in g["pts"]], 2, 3, 3, 1.0, 0.1, 0.1,
                               P.Library(names=P.FULL_NAMES))
    assert np.array_equal(X64, X) and np.array_equal(y64, y)
    # vs the oracle's stencil form: same stencil, only the summation order differs
    Xo, yo = OP.build_dataset_stencil(U, g["pts"], 2, 3, 3, 1.0, 0.1, 0.1, OP.Library(names=OP.FULL_NAMES))
    np.testing.assert_allclose(X, Xo, rtol=1e-10, atol=1e-11)
    with pytest.raises(IndexError):
        P.build_dataset(U, [(1, 3, 3)], 2, 3, 3, 1.0, 0.1, 0.1, P.Library(names=P.FULL_NAMES))
    lib = P.Library(names=P.FULL_NAMES)
    assert np.array_equal(lib.feature_vector(0.5, 1, 2, 3, 4), OP.Library(names=OP.FULL_NAMES).feature_vector(0.5, 1, 2, 3, 4))


def test_patch_stridge_sklearn_dialect(P, golden_patch):
    g = golden_patch
    for X, y, (a, t), ref in zip(g["sk_X"], g["sk_y"], g["sk_grid"], g["sk_coef"]):
        assert_coef_close(P.stridge(X, y, alpha=a, threshold=t), ref, what=f"sklearn dialect {a} {t}")
    # near-constant column far from zero: the shifted statistics keep the variance exact
    rng = np.random.default_rng(4)
    X = rng.standard_normal((120, 8))
    X[:, 0] = 1.0
    X[:, 1] = 1000.0 + 1e-3 * rng.standard_normal(120)
    X[:, 5] = 0.7
    y = X @ rng.standard_normal(8) + 0.01 * rng.standard_normal(120)
    assert_coef_close(P.stridge(X, y, alpha=0.01, threshold=1e-5), OP.stridge(X, y, alpha=0.01, threshold=1e-5),
                      rtol=1e-8, what="ill-centred columns")


def test_patch_ensemble_loop(P, golden_patch):
    """main()'s per-patch loop (patch:395-443) as batched launches == the reference's loop."""
    g = golden_patch
    out = P.fit_patches(g["U"], patch=11, overlap=5, samples_per_patch=40, seed=0)
    assert np.array_equal(out["train_pts"], g["loop_train_pts"])
    assert np.array_equal(out["C"] != 0, g["loop_C"] != 0)
    # north star: coefficients within 1e-8 relative (the reformulations' own floor against the reference is 4.5e-12:
    # tools/patch_floor.py)
    np.testing.assert_allclose(out["C"], g["loop_C"], rtol=1e-8, atol=0)
    assert np.array_equal(out["freq"], g["loop_freq"])
    assert np.array_equal(out["sign_stability"], g["loop_sign_stability"])
    for k in ("median", "q25", "q75", "agg"):
        np.testing.assert_allclose(out[k], g[f"loop_{k}"], rtol=1e-8, atol=0)
    # batched == one-at-a-time drop-in calls
    lib = P.Library(names=P.FULL_NAMES)
    for b in (0, len(out["C"]) - 1):
        X, y = P.build_dataset(g["U"], out["train_pts"][b], 2, 3, 3, 1.0, 0.1, 0.1, lib)
        assert_coef_close(P.stridge(X, y, alpha=0.01, threshold=1e-5), out["C"][b], rtol=1e-9, what=f"patch {b}")


def test_patch_ensemble_larger_vs_oracle(P):
    """A laser-image-shaped stack (config C3 scaled down): every patch's support and coefficients
    against the oracle's per-patch loop."""
    U = synthetic_stack((24, 64, 80), seed=1)
    out = P.fit_patches(U, seed=3)
    ref = OP.run_patches(U, seed=3)
    assert out["C"].shape == ref["C"].shape and len(out["C"]) == 4 * 6
    assert np.array_equal(out["C"] != 0, ref["C"] != 0)
    np.testing.assert_allclose(out["C"], ref["C"], rtol=1e-8, atol=0)
    np.testing.assert_allclose(out["agg"], ref["agg"], rtol=1e-8, atol=0)


def test_fit_metrics_on_gpu(P):
    """rmse / r2_score (ks2d:29-40) and regression_metrics (patch:47-65) from the two-pass reduction kernel."""
    from oracle import ks2d as OK
    from pde_b200 import ks2d as K

    rng = np.random.default_rng(8)
    for n in (1, 7, 1000, 300_001):
        y = 3.0 + rng.standard_normal(n)
        yh = y + 0.1 * rng.standard_normal(n)
        np.testing.assert_allclose(K.rmse(y, yh), OK.rmse(y, yh), rtol=1e-12)
        if n > 1:
            np.testing.assert_allclose(K.r2_score(y, yh), OK.r2_score(y, yh), rtol=1e-11)
            ref, got = OP.regression_metrics(y, yh), P.regression_metrics(y, yh)
            assert set(ref) == set(got)
            for k in ref:
                np.testing.assert_allclose(got[k], ref[k], rtol=1e-9, atol=1e-13, err_msg=k)
    assert K.r2_score(np.arange(5.0), np.arange(5.0)) == 1.0
    X = rng.standard_normal((6, 3))
    assert np.array_equal(K.standardize_transform(X, X.mean(0), X.std(0)), OK.standardize_transform(X, X.mean(0), X.std(0)))


def test_patch_metrics_global_checks_and_gaussian_filter(P, golden_patch):
    """patch:425-429 per-patch train / test regression_metrics, patch:446-465 global held-out test and one-step check
    (continuing the loop's RNG stream), patch:335,343 scipy gaussian_filter: all against outputs of the reference."""
    g = golden_patch
    keys = ("r2", "rmse", "mae", "nrmse", "corr", "resid_mean", "resid_std", "resid_med_abs")
    out = P.fit_patches(g["U"], patch=11, overlap=5, samples_per_patch=40, seed=0)
    assert np.array_equal(out["test_pts"], g["loop_test_pts"])
    for tag in ("train", "test"):
        got = np.stack([out[f"{tag}_metrics"][k] for k in keys], axis=1)
        np.testing.assert_allclose(got, g[f"loop_{tag}_metrics"], rtol=2e-7, atol=1e-10)      # metrics of a 1e-8 coefficient match
    chk = P.global_checks(g["U"], out["agg"], out["rng"], n_global=90, n_step=130)
    assert np.array_equal(np.array(chk["global_points"]), g["global_pts"]) and np.array_equal(np.array(chk["step_points"]), g["step_pts"])
    np.testing.assert_allclose([chk["test_metrics"][k] for k in keys], g["global_test_metrics"], rtol=1e-7, atol=1e-10)
    np.testing.assert_allclose(chk["one_step_rmse"], g["one_step_rmse"][0], rtol=1e-8)
    raw = g["gf_raw"]
    for tag, sigma, src in (("gf_f32_1.0", 1.0, raw), ("gf_f32_1.2", 1.2, raw), ("gf_f64_1.5", 1.5, raw.astype(np.float64))):
        got = P.gaussian_filter(src, sigma)
        assert got.dtype == g[tag].dtype and np.array_equal(got, g[tag]), tag          # bit-identical to scipy
    assert np.array_equal(P.gaussian_filter(raw[0], 1.0), g["gf_f32_1.0"][0])
    # a radius larger than the frame: scipy reflects repeatedly
    from scipy.ndimage import gaussian_filter

    small = np.random.default_rng(2).random((2, 5, 4))
    assert np.array_equal(P.gaussian_filter(small, 2.0), np.array([gaussian_filter(f, sigma=2.0) for f in small]))


def test_fused_gaussian_filter_equals_two_passes_and_scipy():
    """pg_reflect_gauss2d (both axes through shared memory in one pass) against two pg_reflect_conv passes and against
    scipy.ndimage.gaussian_filter itself (patch:335,343): bit-identical for float32 and float64, on frames with ragged
    32 x 128 tiles, frames smaller than the radius (repeated reflection) and radii up to the fused limit."""
    from scipy.ndimage import gaussian_filter

    from pde_b200 import ops

    rng = np.random.default_rng(11)
    for shape, sigmas in (((3, 70, 300), (0.5, 0.75, 1.0, 1.2, 1.5, 2.0, 3.0)), ((2, 33, 129), (1.0, 1.5, 8.0)), ((2, 5, 4), (1.0, 2.0)),
                          ((1, 32, 128), (1.5,)), ((2, 1, 7), (1.0,)), ((1, 64, 257), (9.0,))):
        for dtype in (np.float32, np.float64):
            U = rng.standard_normal(shape).astype(dtype)
            for sigma in sigmas:
                fused = ops.gaussian_filter_frames(U, sigma).cpu().numpy()
                two = ops.gaussian_filter_frames(U, sigma, two_pass=True).cpu().numpy()
                ref = np.array([gaussian_filter(f, sigma=sigma) for f in U])
                assert fused.dtype == ref.dtype
                assert np.array_equal(fused, two), (shape, dtype, sigma)
                assert np.array_equal(fused, ref), (shape, dtype, sigma)

