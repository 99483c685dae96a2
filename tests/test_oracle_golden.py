"""Pin the oracle (oracle/*.py) to the reference: golden vectors made by
tests/golden/make_golden.py from the unmodified reference, the stored outputs of the
reference's notebook 03, and the installed scikit-learn for the patch dialect.  CPU only."""

import numpy as np
import pytest

from oracle import basic, gram, ks2d, patch

RTOL = 1e-12  # oracle vs reference: same NumPy ops, only summation order may differ


def close(a, b, rtol=RTOL, atol=0.0):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


# --------------------------------------------------------------------------- ks2d
def test_ks_stencils_bitexact(golden_ks2d):
    g = golden_ks2d
    U, dx, dy = g["U"], float(g["dx"]), float(g["dy"])
    gx, gy = ks2d.gradients(U[3], dx, dy)
    assert np.array_equal(gx, g["gx3"]) and np.array_equal(gy, g["gy3"])
    assert np.array_equal(ks2d.laplacian(U[3], dx, dy), g["lap3"])


def test_ks_dictionaries_bitexact(golden_ks2d):
    g = golden_ks2d
    U, dx, dy = g["U"], float(g["dx"]), float(g["dy"])
    names, terms = ks2d.build_dictionary(U[:-1], dx, dy)
    assert names == list(g["names_rich"]) == ks2d.RICH_NAMES
    assert np.array_equal(np.stack([terms[n] for n in names]), g["rich_terms"])
    names, terms = ks2d.build_dictionary_true(U[:-1], dx, dy, include_advection=True)
    assert names == list(g["names_adv"])
    assert np.array_equal(np.stack([terms[n] for n in names]), g["adv_terms"])
    assert ks2d.build_dictionary_true(U[:-1], dx, dy)[0] == list(g["names_true"]) == ks2d.TRUE_NAMES


@pytest.mark.parametrize("tag,block", [("388", (3, 8, 8)), ("453", (4, 5, 3)), ("111", (1, 1, 1))])
def test_ks_blockwise(golden_ks2d, tag, block):
    g = golden_ks2d
    U, dx, dy, DT = g["U"], float(g["dx"]), float(g["dy"]), float(g["DT"])
    Ut = (U[1:] - U[:-1]) / DT
    names, terms = ks2d.build_dictionary(U[:-1], dx, dy)
    X, y = ks2d.build_blockwise_dataset(Ut, terms, names, block_t=block[0], block_x=block[1], block_y=block[2])
    assert X.shape == g[f"bw{tag}_X_rich"].shape
    close(X, g[f"bw{tag}_X_rich"], rtol=1e-11, atol=1e-13)
    close(y, g[f"bw{tag}_y"], rtol=1e-11, atol=1e-13)
    Xl, yl = ks2d.build_blockwise_dataset_loops(Ut, terms, names, block_t=block[0], block_x=block[1], block_y=block[2])
    assert np.array_equal(Xl, g[f"bw{tag}_X_rich"]) and np.array_equal(yl, g[f"bw{tag}_y"])


def test_ks_blockwise_errors_and_nonfinite():
    Ut = np.ones((4, 4, 4))
    with pytest.raises(ValueError):
        ks2d.build_blockwise_dataset(Ut[0], {"a": Ut}, ["a"], block_t=1, block_x=1, block_y=1)
    with pytest.raises(ValueError):
        ks2d.build_blockwise_dataset(Ut, {"a": Ut}, ["a"], block_t=0, block_x=1, block_y=1)
    a = Ut.copy()
    a[0, 0, 0] = np.nan
    X, y = ks2d.build_blockwise_dataset(Ut, {"a": a}, ["a"], block_t=2, block_x=2, block_y=2)
    assert X.shape == (7, 1) and y.shape == (7,)


def test_ks_stridge(golden_ks2d):
    g = golden_ks2d
    m, s = ks2d.standardize_fit(g["bw111_X_rich"])
    close(m, g["std_mean"]), close(s, g["std_scale"])
    close(ks2d.ridge_fit(ks2d.standardize_transform(g["bw111_X_rich"], m, s), g["bw111_y"], 1e-3), g["ridge_fit"], 1e-10)
    for X, key in [(g["bw453_X_rich"], "stridge_rich"), (g["bw111_X_rich"], "stridge_rich_pointwise"),
                   (g["bw111_X_true"], "stridge_true_pointwise")]:
        y = g["bw453_y"] if key == "stridge_rich" else g["bw111_y"]
        for (a, t), ref in zip(g["stridge_grid"], g[key]):
            c = ks2d.stridge(X, y, alpha=a, threshold=t, max_iter=25)
            assert np.array_equal(c != 0, ref != 0)
            close(c, ref, rtol=1e-9)


def test_ks_stridge_from_stats_matches_rows(golden_ks2d):
    """The Theta-free formulation reproduces the row-form support bit for bit."""
    g = golden_ks2d
    for X, y, key, const in [(g["bw453_X_rich"], g["bw453_y"], "stridge_rich", (0,)),
                             (g["bw111_X_rich"], g["bw111_y"], "stridge_rich_pointwise", (0,)),
                             (g["bw111_X_true"], g["bw111_y"], "stridge_true_pointwise", ())]:
        s = gram.pack_stats(X, y)
        p = X.shape[1]
        for (a, t), ref in zip(g["stridge_grid"], g[key]):
            c = gram.stridge_from_stats(s, p, dialect=gram.DIALECT_KS, alpha=a, threshold=t, max_iter=25,
                                        const_cols=const)
            assert np.array_equal(c != 0, ref != 0), (key, a, t)
            close(c, ref, rtol=1e-8)


@pytest.mark.slow
def test_ks_simulate_matches_reference_checksum(ks_default_stack, golden_configs):
    U, dx, dy, DT = ks_default_stack
    chk = golden_configs["full_precision"]["c1"]["U_checksum"]
    assert U.shape == (2000, 100, 100) and (dx, dy, DT) == (0.5, 0.5, 1e-3)
    assert [float(U.sum()), float((U ** 2).sum()), float(U[-1, 1, 2])] == chk


@pytest.mark.slow
@pytest.mark.parametrize("tag,kw", [
    ("c1", dict(method="pointwise", dictionary="true")),
    ("c2", dict(method="blockwise", dictionary="true")),
    ("c2_rich_sweep", dict(method="blockwise", dictionary="rich", grid_search=True)),
])
def test_ks_configs_match_reference_main(ks_default_stack, golden_configs, tag, kw):
    """Configs C1 / C2 / C2+rich+sweep: the oracle pipeline vs the numbers main() printed."""
    U, dx, dy, DT = ks_default_stack
    if tag != "c1":
        U = ks2d.add_noise(U, 0.05, seed=999)
    out = ks2d.run_config(U, dx, dy, DT, **kw)
    gold = golden_configs[tag]
    full = golden_configs["full_precision"][tag]
    assert list(out["X_shape"]) == gold["X_shape"] == full["X_shape"]
    close(out["X"][:4], np.array(full["X_head"]), rtol=1e-11, atol=1e-14)
    hyper = gold["hyper"]
    assert out["alpha"] == hyper["alpha"] and out["threshold"] == hyper["threshold"]
    assert out["n_active"] == hyper["n_active"]
    close(out["r2_test"], hyper["r2_test"], rtol=1e-8)
    close(out["rmse_test"], hyper["rmse_test"], rtol=1e-6)
    for n, c in zip(out["names"], out["coeffs"]):
        if n in gold["coeffs_printed"]:
            assert abs(c - gold["coeffs_printed"][n]) < 1.5e-6
    if "table" in out:
        for (a, thr, r2, err, na), row in zip(out["table"], full["table"]):
            assert (a, thr, na) == (row["alpha"], row["threshold"], row["n_active"])
            close(r2, row["r2_test"], rtol=1e-7)


@pytest.mark.slow
def test_notebook03_known_answers():
    """notebooks/03_synthetic_benchmark_verification.ipynb cells 0-4 stored outputs."""
    U, dx, dy, DT = ks2d.simulate(ks2d.SimConfig(Nx=64, Ny=64, n_seconds=0.5, seed=42))
    assert U.shape == (500, 64, 64)
    assert (float(U.min()), float(U.max())) == (-0.09916768934803566, 0.10317702881067896)
    names, X, y, _ = ks2d.make_dataset(U, dx, dy, DT, method="pointwise", dictionary="true", n_sample=50_000)
    assert X.shape == (50000, 3)
    assert (float(y.mean()), float(y.std())) == (0.0005824567648833928, 0.37341971631854826)
    perm = np.random.default_rng(1).permutation(len(y))
    tr = perm[: int(0.7 * len(y))]
    c = ks2d.stridge(X[tr], y[tr], alpha=1e-6, threshold=1e-10)
    close(c, [-1.0, -1.0, -0.5], rtol=1e-7)
    close(ks2d.rmse(y[tr], X[tr] @ c), 3.216230714271724e-11, rtol=0.05)
    # same fit through the statistics-only formulation
    c2 = gram.stridge_from_stats(gram.pack_stats(X[tr], y[tr]), 3, dialect=gram.DIALECT_KS, alpha=1e-6,
                                 threshold=1e-10, max_iter=25)
    close(c2, c, rtol=1e-8)


@pytest.mark.slow
def test_ks_sweep_from_stats_selects_same_model(ks_default_stack, golden_configs):
    """C2+rich+sweep: the statistics-only sweep picks the reference's (alpha, thr, support)."""
    U, dx, dy, DT = ks_default_stack
    U = ks2d.add_noise(U, 0.05, seed=999)
    names, X, y, rng = ks2d.make_dataset(U, dx, dy, DT, method="blockwise", dictionary="rich")
    tr, te, _ = ks2d.split_and_scale(names, X, y, rng)
    best = gram.ks_fit_from_stats(gram.pack_stats(X[tr], y[tr]), gram.pack_stats(X[te], y[te]), 9,
                                  alphas=ks2d.GRID_ALPHAS, thresholds=ks2d.GRID_THRESHOLDS, const_cols=(0,))
    full = golden_configs["full_precision"]["c2_rich_sweep"]["table"]
    for (a, thr, r2, err, na), row in zip(best["table"], full):
        assert na == row["n_active"]
        close(r2, row["r2_test"], rtol=1e-7)
    hyper = golden_configs["c2_rich_sweep"]["hyper"]
    assert (best["alpha"], best["threshold"], best["n_active"]) == (hyper["alpha"], hyper["threshold"], hyper["n_active"])


# --------------------------------------------------------------------------- basic_usage
def test_basic_derivatives_and_library(golden_basic):
    g = golden_basic
    ut, u, ux, uy, lap = basic.compute_derivatives(g["small_u"], *g["small_d"])
    for a, k in [(ut, "small_ut"), (ux, "small_ux"), (uy, "small_uy"), (lap, "small_lap")]:
        assert np.array_equal(a, g[k]), k
    Theta, names = basic.build_library(u, ux, uy, lap)
    assert names == basic.TERM_NAMES and np.array_equal(Theta, g["small_Theta"])


def test_basic_stridge(golden_basic):
    g = golden_basic
    ut = basic.compute_derivatives(g["small_u"], *g["small_d"])[0].reshape(-1)
    Th = g["small_Theta"]
    s = gram.pack_stats(Th, ut)
    for (a, t), ref in zip(g["small_grid"], g["small_coef"]):
        c = basic.stridge_regression(Th, ut, alpha=a, threshold=t)
        assert np.array_equal(c != 0, ref != 0)
        close(c, ref, rtol=1e-10)
        c2 = gram.stridge_from_stats(s, 6, dialect=gram.DIALECT_BASIC, alpha=a, threshold=t, max_iter=10)
        assert np.array_equal(c2 != 0, ref != 0)
        close(c2, ref, rtol=1e-8)
    assert np.array_equal(basic.stridge_regression(Th, ut, max_iter=0), g["small_coef_iter0"])


def test_basic_default_example(golden_basic):
    g = golden_basic
    u, x, y, t = basic.generate_synthetic_data(n_frames=30, h=60, w=60)
    assert np.array_equal(u[::7, ::11, ::13], g["default_u_sample"])
    d = (x[1] - x[0], y[1] - y[0], t[1] - t[0])
    assert np.array_equal(np.array(d), g["default_spacing"])
    ut, uu, ux, uy, lap = basic.compute_derivatives(u, *d)
    Theta, _ = basic.build_library(uu, ux, uy, lap)
    assert Theta.shape == tuple(g["default_Theta_shape"]) == (90944, 6)
    close(Theta.T @ Theta, g["default_G"], rtol=1e-13)
    c = basic.stridge_regression(Theta, ut.reshape(-1))
    close(c, g["default_coef"], rtol=1e-10)
    close(c, [0, -0.0256799572, -0.4924756608, -0.2951039536, 0.0512370791, 0], rtol=1e-8)


# --------------------------------------------------------------------------- patch
def test_patch_poly_derivatives(golden_patch):
    g = golden_patch
    U = g["U"]
    assert U.dtype == np.float32
    d = np.array([patch.local_poly_derivatives(U, *p, 2, 3, 3, 1.0, 0.1, 0.1) for p in g["pts"]])
    close(d, g["derivs"], rtol=1e-9, atol=1e-11)
    d2 = np.array([patch.local_poly_derivatives(U, *p, 1, 2, 2, 0.5, 0.2, 0.3) for p in g["pts"]])
    close(d2, g["derivs_deg2_r1"], rtol=1e-9, atol=1e-11)
    # the fixed-stencil form (what the GPU computes)
    W = patch.poly_stencil(2, 3, 3, 1.0, 0.1, 0.1)
    assert W.shape == (6, 245)
    lib = patch.Library(names=patch.FULL_NAMES)
    X, y = patch.build_dataset_stencil(U, g["pts"], 2, 3, 3, 1.0, 0.1, 0.1, lib, W)
    close(X, g["X8"], rtol=1e-8, atol=1e-9)
    close(y, g["y8"], rtol=1e-8, atol=1e-10)
    X6, _ = patch.build_dataset_stencil(U, g["pts"], 2, 3, 3, 1.0, 0.1, 0.1, patch.Library(names=patch.MODEL4_NAMES), W)
    close(X6, g["X6"], rtol=1e-8, atol=1e-9)


def test_patch_grid(golden_patch):
    g = golden_patch
    assert np.array_equal(np.array(patch.patch_grid(30, 34, 9, 4)), g["patch_grid_30_34_9_4"])
    assert len(patch.patch_grid(1024, 1024, 21, 10)) == int(g["patch_grid_1024"][0]) == 8464


def test_patch_stridge_vs_reference_and_sklearn(golden_patch):
    g = golden_patch
    from sklearn.linear_model import Ridge
    from sklearn.preprocessing import StandardScaler

    for X, y, (a, t), ref in zip(g["sk_X"], g["sk_y"], g["sk_grid"], g["sk_coef"]):
        c = patch.stridge(X, y, alpha=a, threshold=t)
        assert np.array_equal(c != 0, ref != 0)
        close(c, ref, rtol=1e-9, atol=1e-15)
        c2 = gram.stridge_from_stats(gram.pack_stats(X, y), X.shape[1], dialect=gram.DIALECT_SKLEARN, alpha=a,
                                     threshold=t, max_iter=25)
        assert np.array_equal(c2 != 0, ref != 0)
        close(c2, ref, rtol=1e-8, atol=1e-15)
        sc = StandardScaler().fit(X)
        m, s, _ = patch._scaler_fit(X)
        close(m, sc.mean_), close(s, sc.scale_)
        close(patch._ridge_intercept_coef((X - m) / s, y, a), Ridge(alpha=a).fit(sc.transform(X), y).coef_, rtol=1e-9, atol=1e-14)


def test_patch_loop(golden_patch):
    """The per-patch loop of main() (sampling order, datasets, STRidge, aggregation)."""
    g = golden_patch
    U = g["U"]
    out = patch.run_patches(U, patch=11, overlap=5, samples_per_patch=40, seed=0)
    assert np.array_equal(np.stack([s[0] for s in out["samples"]]), g["loop_train_pts"])
    assert np.array_equal(np.stack([s[1] for s in out["samples"]]), g["loop_test_pts"])
    assert np.array_equal(out["C"] != 0, g["loop_C"] != 0)
    close(out["C"], g["loop_C"], rtol=1e-6, atol=1e-9)
    for k in ("freq", "sign_stability"):
        assert np.array_equal(out[k], g[f"loop_{k}"])
    for k in ("median", "q25", "q75", "agg"):
        close(out[k], g[f"loop_{k}"], rtol=1e-6, atol=1e-9)


# --------------------------------------------------------------------------- sign-constrained STRidge
@pytest.fixture(scope="session")
def golden_signed():
    from conftest import GOLDEN

    return np.load(GOLDEN / "ks2d_signed.npz")


@pytest.mark.parametrize("tag,xkey,ykey", [("true_pw", "bw111_X_true", "bw111_y"), ("rich_pw", "bw111_X_rich", "bw111_y"),
                                           ("rich_453", "bw453_X_rich", "bw453_y")])
def test_sign_constrained_rows_and_stats_match_reference(golden_ks2d, golden_signed, tag, xkey, ykey):
    """oracle.ks2d.stridge_sign_constrained (rows) and its statistics form against the outputs of the
    reference's ks2d:552-600 on the same rows: identical support, coefficients to 1e-9 / 1e-8."""
    from oracle import gram, ks2d

    X, y, g = golden_ks2d[xkey], golden_ks2d[ykey], golden_signed
    p = X.shape[1]
    const = [j for j in range(p) if np.all(X[:, j] == X[0, j])]
    stats = gram.pack_stats(X, y)
    for signs, (a, t), ref in zip(g[f"{tag}_signs"], g[f"{tag}_grid"], g[f"{tag}_coef"]):
        rows = ks2d.stridge_sign_constrained(X, y, alpha=a, threshold=t, max_iter=25, signs=[int(v) for v in signs])
        assert np.array_equal(rows != 0, ref != 0), (tag, signs, a, t)
        np.testing.assert_allclose(rows, ref, rtol=1e-9, atol=0)
        st = gram.stridge_from_stats(stats, p, dialect=gram.DIALECT_KS, alpha=a, threshold=t, max_iter=25,
                                     const_cols=const, signs=[int(v) for v in signs])
        assert np.array_equal(st != 0, ref != 0), (tag, signs, a, t)
        np.testing.assert_allclose(st, ref, rtol=1e-8, atol=0)
    none = ks2d.stridge_sign_constrained(X, y, alpha=1e-3, threshold=1e-6, max_iter=25, signs=None)
    np.testing.assert_allclose(none, g[f"{tag}_none"], rtol=1e-9)
    it0 = ks2d.stridge_sign_constrained(X, y, alpha=1e-3, threshold=1e-6, max_iter=0, signs=[-1] * p)
    np.testing.assert_allclose(it0, g[f"{tag}_iter0"], rtol=1e-9)


# --------------------------------------------------------------------------- rollout check
@pytest.mark.parametrize("tag", ["c1", "c2", "c2_rich_sweep"])
def test_rollout_matches_reference_main(tag, golden_configs, ks_default_stack):
    """oracle.ks2d.rollout_errors against the 50 per-step RMSEs the reference's main() computed (captured at
    full precision by tests/golden/make_golden.py): bit-identical arithmetic."""
    import json

    from conftest import GOLDEN
    from oracle import ks2d

    g = json.loads((GOLDEN / "ks2d_rollout.json").read_text())[tag]
    from helpers import rollout_case

    Uo, dx, dy, DT, names, coef = rollout_case(tag, golden_configs, ks_default_stack)
    errs = ks2d.rollout_errors(Uo, dx, dy, DT, names, coef, 50)
    assert len(errs) == g["n_steps"] == 50
    np.testing.assert_allclose(errs, g["errs"], rtol=1e-12, atol=0)
    np.testing.assert_allclose([errs[0], errs[-1], errs.mean()], g["printed"], rtol=2e-3)


# --------------------------------------------------------------------------- denoising prologue
def test_smoothing_prologue_matches_reference():
    """oracle restatements of ks2d:125-161 bit for bit, and the separable periodic-Gaussian taps the GPU path
    uses (no FFT) against the reference's FFT result."""
    from conftest import GOLDEN
    from oracle import ks2d
    from pde_b200.ops import periodic_gaussian_taps

    g = np.load(GOLDEN / "ks2d_smooth.npz")
    for tag in "abc":
        f = g[f"frame_{tag}"]
        for sig in (0.8, 1.5, 4.0):
            ref = g[f"gauss_{tag}_{sig}"]
            assert np.array_equal(ks2d.gaussian_smooth_periodic_2d(f, sig), ref)
            tmp, out = np.zeros_like(f), np.zeros_like(f)
            for o, w in zip(*periodic_gaussian_taps(f.shape[0], sig)):
                tmp += w * np.roll(f, o, axis=0)
            for o, w in zip(*periodic_gaussian_taps(f.shape[1], sig)):
                out += w * np.roll(tmp, o, axis=1)
            assert np.abs(out - ref).max() <= 1e-14 * np.abs(ref).max(), (tag, sig)
    assert len(periodic_gaussian_taps(100, 4.0)[0]) < 100 and len(periodic_gaussian_taps(100, 0.8)[0]) == 100
    for w in (3, 5, 9):
        assert np.array_equal(ks2d.time_smooth_moving_average(g["stack"], w), g[f"tavg_{w}"])
    with pytest.raises(ValueError):
        ks2d.time_smooth_moving_average(g["stack"], 4)


# --------------------------------------------------------------------------- bootstrap ensemble (ks2d:603-642)
def test_ensemble_stridge_matches_reference(golden_ks2d):
    from conftest import GOLDEN
    from oracle import ks2d

    g = np.load(GOLDEN / "ks2d_ensemble.npz")
    for tag, X in (("true", golden_ks2d["bw111_X_true"]), ("rich", golden_ks2d["bw111_X_rich"])):
        for k in range(2):
            a, t, nb, frac, seed = g[f"{tag}_{k}_args"]
            med, std = ks2d.ensemble_stridge(X, golden_ks2d["bw111_y"], alpha=a, threshold=t, n_bootstrap=int(nb),
                                             subsample_frac=frac, seed=int(seed))
            assert np.array_equal(med, g[f"{tag}_{k}_median"]) and np.array_equal(std, g[f"{tag}_{k}_std"])


# --------------------------------------------------------------------------- analyze_results dialect
def test_analyze_oracle_matches_the_reference_lines():
    """oracle/analyze.py against tests/golden/analyze.npz (the reference's own source lines executed on a synthetic
    stack): derivative slices bit-identical, the six models' coefficients / scales / metrics, one-step and rollout."""
    from conftest import GOLDEN
    from oracle import analyze as OA

    g = np.load(GOLDEN / "analyze.npz")
    dx, dy, dt = g["spacing"]
    d = OA.derivatives(g["U"], dx, dy, dt)
    for k, gk in (("u", "u"), ("u_x", "u_x"), ("u_y", "u_y"), ("u_xx", "u_xx"), ("u_yy", "u_yy"), ("u_t", "u_t"), ("lap", "laplacian")):
        assert np.array_equal(d[k], g["d_" + gk]), k
    assert d["u"].shape == tuple(g["aligned"])
    tr, te = OA.split_time(d["u"].shape[0], 0.7)
    assert tr.stop == int(g["train_stop"][0])
    res = OA.fit_models(g["U"], dx, dy, dt)
    keys = ("r2", "rmse", "mae", "nrmse", "corr", "resid_mean", "resid_std", "resid_med_abs")
    for idx, (name, r) in enumerate(res.items(), start=1):
        assert r["names"] == [str(n) for n in g[f"m{idx}_names"]]
        assert np.array_equal(r["coeffs"] != 0, g[f"m{idx}_coeffs"] != 0)
        np.testing.assert_allclose(r["coeffs"], g[f"m{idx}_coeffs"], rtol=1e-10)
        np.testing.assert_allclose(r["scale"], g[f"m{idx}_scale"], rtol=1e-13)
        np.testing.assert_allclose([r["test"][k] for k in keys], g[f"m{idx}_test_metrics"], rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose([r["train"][k] for k in keys], g[f"m{idx}_train_metrics"], rtol=1e-9, atol=1e-12)
    for idx in (3, 6):
        names = [str(n) for n in g[f"m{idx}_names"]]
        ro = np.array([[OA.rollout_k_rmse(d["u"], names, g[f"m{idx}_coeffs"], k, sl, dx, dy, dt)[q] for q in ("rmse", "nrmse")]
                       for k in (1, 3) for sl in (tr, te)])
        np.testing.assert_allclose(ro, g[f"m{idx}_rollout"], rtol=1e-12)
    X, y = OA.rows(d, OA.FULL_NAMES, tr)
    for (a, th), ref in zip(g["m6_grid"], g["m6_grid_coeffs"]):
        c, _ = OA.stridge(X, y, alpha=a, threshold=th)
        assert np.array_equal(c != 0, ref != 0)
        np.testing.assert_allclose(c, ref, rtol=1e-9)


def test_analyze_statistics_formulation_selects_the_same_support():
    """The sklearn-dialect STRidge on sufficient statistics (what K3 runs) equals the row form for every model."""
    from conftest import GOLDEN
    from oracle import analyze as OA

    g = np.load(GOLDEN / "analyze.npz")
    dx, dy, dt = g["spacing"]
    d = OA.derivatives(g["U"], dx, dy, dt)
    tr, _ = OA.split_time(d["u"].shape[0], 0.7)
    for idx, (name, names) in enumerate(OA.MODELS.items(), start=1):
        X, y = OA.rows(d, names, tr)
        s = gram.pack_stats(X, y)
        c = gram.stridge_from_stats(s, len(names), dialect=1, alpha=0.01, threshold=1e-5, max_iter=20, const_cols=(0,))
        scale = g[f"m{idx}_scale"]
        c = c * (scale + 1e-12) / scale            # the patch dialect's unscaling adds 1e-12 (patch:98); ar:578 does not
        assert np.array_equal(c != 0, g[f"m{idx}_coeffs"] != 0), name
        np.testing.assert_allclose(c, g[f"m{idx}_coeffs"], rtol=1e-8)


# --------------------------------------------------------------------------- patch_based_sindy
def test_sindy_oracle_matches_the_reference_class():
    """oracle/sindy.py against tests/golden/sindy.npz (the unmodified PatchBasedSINDy class on synthetic images), the
    scrambled feature view included."""
    from conftest import GOLDEN
    from oracle import sindy as OS

    g = np.load(GOLDEN / "sindy.npz")
    dt, dx, dy, ps, ov = g["params"]
    ps, ov = int(ps), int(ov)
    assert tuple(g["lib_shape"]) == (ps, 11 * ps)                       # column_stack of 2-D arrays concatenates columns
    seq = [f[24:56, 48:80].copy() for f in g["images"]]
    X, y = OS.patch_rows(seq, dx, dy, dt)
    assert np.array_equal(X.reshape(6, 5, 5, 11)[0], g["lib_view_sample"])
    c, q = OS.fit_rows(X, y, 0.01)
    np.testing.assert_allclose(c, g["one_coeffs"], rtol=1e-10)
    np.testing.assert_allclose(q, g["one_quality"][0], rtol=1e-10)
    ens, info = OS.ensemble(g["images"], ps, ov, dx, dy, dt, min_patches=3)
    np.testing.assert_allclose(info["patch_coeffs"], g["ens_patch_coeffs"], rtol=1e-10)
    np.testing.assert_allclose(info["patch_qualities"], g["ens_patch_qualities"], rtol=1e-10, atol=1e-15)
    assert np.array_equal(ens != 0, g["ens_coeffs"] != 0)
    np.testing.assert_allclose(ens, g["ens_coeffs"], rtol=1e-10)
    # the unscrambled variant differs (it is what the code presumably meant, not what it does)
    Xu, _ = OS.patch_rows(seq, dx, dy, dt, scramble=False)
    assert not np.array_equal(Xu, X) and np.array_equal(Xu[:, 0], np.ones(len(Xu)))
