"""GPU parity, KS-2D dialect: every test calls the CUDA path through the C ABI (pde_b200 ->
ctypes -> libpdegram.so) and checks it against the golden fixtures / the oracle on the same
seeded inputs.  Bars: bit-exact for the materialised stencil terms, GRAM_RTOL for statistics,
COEF_RTOL + identical support for coefficients."""

import numpy as np
import pytest

from helpers import COEF_RTOL, GRAM_RTOL, assert_coef_close, assert_stats_close, ks_rows
from oracle import gram, ks2d as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import pde_b200
    from pde_b200 import ks2d

    pde_b200.load()
    return ks2d


@pytest.fixture(scope="module")
def L():
    from pde_b200 import _lib

    return _lib


@pytest.fixture(scope="module")
def ops():
    from pde_b200 import ops

    return ops


def test_stencils_bitexact(K, golden_ks2d):
    g = golden_ks2d
    U, dx, dy = g["U"], float(g["dx"]), float(g["dy"])
    gx, gy = K.gradients(U[3], dx, dy)
    assert np.array_equal(gx, g["gx3"]) and np.array_equal(gy, g["gy3"])
    assert np.array_equal(K.laplacian(U[3], dx, dy), g["lap3"])


def test_dictionaries_bitexact(K, golden_ks2d):
    g = golden_ks2d
    U, dx, dy = g["U"], float(g["dx"]), float(g["dy"])
    names, terms = K.build_dictionary(U[:-1], dx, dy)
    assert names == list(g["names_rich"])
    for k, n in enumerate(names):
        assert terms[n].shape == U[:-1].shape and terms[n].dtype == np.float64
        assert np.array_equal(terms[n], g["rich_terms"][k]), n
    names, terms = K.build_dictionary_true(U[:-1], dx, dy, include_advection=True)
    assert names == list(g["names_adv"])
    for k, n in enumerate(names):
        assert np.array_equal(terms[n], g["adv_terms"][k]), n
    assert K.build_dictionary_true(U[:-1], dx, dy)[0] == list(g["names_true"])
    with pytest.raises(NotImplementedError):
        K.build_dictionary(U[:-1], dx, dy, deriv="spectral")


@pytest.mark.parametrize("shape", [(3, 1, 1), (2, 2, 3), (2, 5, 2), (4, 3, 4)])
def test_stencils_tiny_periodic_shapes(K, shape):
    """np.roll wraps any size, including extents 1 and 2 where +-1 / +-2 alias each other."""
    U = np.random.default_rng(5).standard_normal(shape)
    names, terms = K.build_dictionary(U, 0.3, 0.7)
    _, ref = O.build_dictionary(U, 0.3, 0.7)
    for n in names:
        assert np.array_equal(terms[n], ref[n]), n


@pytest.mark.parametrize("tag,block", [("388", (3, 8, 8)), ("453", (4, 5, 3)), ("111", (1, 1, 1))])
def test_build_blockwise_dataset(K, golden_ks2d, tag, block):
    g = golden_ks2d
    U, dx, dy, DT = g["U"], float(g["dx"]), float(g["dy"]), float(g["DT"])
    Ut = (U[1:] - U[:-1]) / DT
    names, terms = O.build_dictionary(U[:-1], dx, dy)
    X, y = K.build_blockwise_dataset(Ut, terms, names, block_t=block[0], block_x=block[1], block_y=block[2])
    assert X.shape == g[f"bw{tag}_X_rich"].shape and y.shape == g[f"bw{tag}_y"].shape
    np.testing.assert_allclose(X, g[f"bw{tag}_X_rich"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(y, g[f"bw{tag}_y"], rtol=1e-11, atol=1e-13)


def test_build_blockwise_dataset_errors_and_nonfinite(K):
    Ut = np.ones((4, 4, 4))
    with pytest.raises(ValueError, match="Ut must be"):
        K.build_blockwise_dataset(Ut[0], {"a": Ut}, ["a"], block_t=1, block_x=1, block_y=1)
    with pytest.raises(ValueError, match="block sizes"):
        K.build_blockwise_dataset(Ut, {"a": Ut}, ["a"], block_t=0, block_x=1, block_y=1)
    a = Ut.copy()
    a[0, 0, 0] = np.nan
    X, y = K.build_blockwise_dataset(Ut, {"a": a}, ["a"], block_t=2, block_x=2, block_y=2)
    assert X.shape == (7, 1) and y.shape == (7,)
    X, y = K.build_blockwise_dataset(Ut * np.nan, {"a": a}, ["a"], block_t=2, block_x=2, block_y=2)
    assert X.shape == (0, 1) and y.shape == (0,)


def test_stridge_golden(K, golden_ks2d):
    g = golden_ks2d
    for X, y, key in [(g["bw453_X_rich"], g["bw453_y"], "stridge_rich"),
                      (g["bw111_X_rich"], g["bw111_y"], "stridge_rich_pointwise"),
                      (g["bw111_X_true"], g["bw111_y"], "stridge_true_pointwise")]:
        for (a, t), ref in zip(g["stridge_grid"], g[key]):
            c = K.stridge(X, y, alpha=a, threshold=t, max_iter=25)
            assert_coef_close(c, ref, what=f"{key} alpha={a} thr={t}")
    m, s = K.standardize_fit(g["bw111_X_rich"])
    np.testing.assert_allclose(m, g["std_mean"], rtol=1e-10, atol=1e-15)
    np.testing.assert_allclose(s, g["std_scale"], rtol=1e-10)
    Xs = O.standardize_transform(g["bw111_X_rich"], g["std_mean"], g["std_scale"])
    np.testing.assert_allclose(K.ridge_fit(Xs, g["bw111_y"], 1e-3), g["ridge_fit"], rtol=COEF_RTOL)


def test_stridge_max_iter_edge_cases(K, golden_ks2d):
    g = golden_ks2d
    X, y = g["bw453_X_rich"], g["bw453_y"]
    for mi in (0, 1, 2):
        for a, t in [(1e-3, 1e-6), (1e-2, 0.2), (1e-1, 5.0)]:
            assert_coef_close(K.stridge(X, y, alpha=a, threshold=t, max_iter=mi),
                              O.stridge(X, y, alpha=a, threshold=t, max_iter=mi), what=f"max_iter={mi} {a} {t}")


LIBS = [("true", False, "LIB_KS_TRUE"), ("true", True, "LIB_KS_TRUE_ADV"), ("rich", False, "LIB_KS_RICH")]


@pytest.mark.parametrize("dictionary,adv,libname", LIBS)
@pytest.mark.parametrize("block", [(1, 1, 1), (3, 8, 8), (4, 5, 3), (2, 10, 6), (50, 7, 5)])
def test_fused_gram_generic_vs_oracle(ops, L, golden_ks2d, dictionary, adv, libname, block):
    """K1 (generic kernel): statistics of the fused path == statistics of the oracle's rows."""
    g = golden_ks2d
    U, dx, dy, DT = g["U"], float(g["dx"]), float(g["dy"]), float(g["DT"])
    names, X, y = ks_rows(U, dx, dy, DT, dictionary, adv, block)
    stats = ops.fd_lib_gram(U, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=getattr(L, libname), block=block,
                            variant=L.VARIANT_GENERIC).cpu().numpy()
    assert_stats_close(stats[0], gram.pack_stats(X, y), len(names))


def test_fused_gram_folds(ops, L, golden_ks2d):
    g = golden_ks2d
    U, dx, dy, DT = g["U"], float(g["dx"]), float(g["dy"]), float(g["DT"])
    names, X, y = ks_rows(U, dx, dy, DT, "rich", False, (3, 8, 8))
    fold = np.random.default_rng(2).integers(0, 3, size=len(y)).astype(np.uint8)
    stats = ops.fd_lib_gram(U, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_RICH, block=(3, 8, 8),
                            fold_of_row=fold, n_folds=3, variant=L.VARIANT_GENERIC).cpu().numpy()
    for f in range(3):
        assert_stats_close(stats[f], gram.pack_stats(X[fold == f], y[fold == f]), 9)
    # time-holdout folds: a block takes the fold of its first frame
    T1 = U.shape[0] - 1
    fof = (np.arange(T1) >= 9).astype(np.int32)
    stats = ops.fd_lib_gram(U, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 4, 4),
                            fold_of_frame=fof, n_folds=2, variant=L.VARIANT_GENERIC).cpu().numpy()
    names, X, y = ks_rows(U, dx, dy, DT, "true", False, (3, 4, 4))
    rows_per_tb = len(y) // 5
    row_fold = np.repeat(fof[::3], rows_per_tb)
    for f in range(2):
        assert_stats_close(stats[f], gram.pack_stats(X[row_fold == f], y[row_fold == f]), 3)
    with pytest.raises(ValueError, match="fold_of_row"):
        ops.fd_lib_gram(U, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8),
                        fold_of_row=fold[:-1], n_folds=3)


def test_fused_gram_edge_shapes(ops, L):
    """Single frame (no u_t -> zero rows), 2 frames, extents smaller than the stencil radius."""
    rng = np.random.default_rng(9)
    s = ops.fd_lib_gram(rng.standard_normal((1, 6, 6)), 1.0, 1.0, 1.0, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_RICH).cpu().numpy()
    assert not s.any()
    for shape in [(2, 1, 1), (2, 2, 2), (3, 1, 7), (3, 4, 1)]:
        U = rng.standard_normal(shape)
        names, X, y = ks_rows(U, 0.5, 0.25, 0.1, "rich", False, (1, 1, 1))
        s = ops.fd_lib_gram(U, 0.5, 0.25, 0.1, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_RICH).cpu().numpy()
        assert_stats_close(s[0], gram.pack_stats(X, y), 9)


def test_fused_gram_nonfinite_rows_are_dropped(ops, L, golden_ks2d):
    g = golden_ks2d
    U, dx, dy, DT = g["U"].copy(), float(g["dx"]), float(g["dy"]), float(g["DT"])
    U[4, 7, 3] = np.nan
    U[9, 0, 0] = np.inf
    names, X, y = ks_rows(U, dx, dy, DT, "true", False, (3, 8, 8))
    stats, bad = ops.fd_lib_gram(U, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8),
                                 variant=L.VARIANT_GENERIC, return_nonfinite=True)
    n_all = 5 * 3 * 2
    assert int(bad[0].item()) == n_all - len(y) > 0
    assert_stats_close(stats.cpu().numpy()[0], gram.pack_stats(X, y), 3)


def test_c_abi_error_codes(L):
    """Bad arguments come back as PG_EINVAL with a message, never as a crash."""
    import torch

    lib = L.load()
    U = torch.zeros((3, 8, 8), dtype=torch.float64, device="cuda")
    out = torch.zeros(64, dtype=torch.float64, device="cuda")
    rc = lib.pg_fd_lib_gram(U.data_ptr(), 3, 8, 8, 1.0, 1.0, 1.0, 0, 0, 0, 1, 1, None, None, 1, out.data_ptr(), None, 0, None)
    assert rc == -1 and b"block sizes" in lib.pg_last_error()
    rc = lib.pg_fd_lib_gram(U.data_ptr(), 3, 8, 8, 1.0, 1.0, 1.0, 0, L.LIB_BASIC, 1, 1, 1, None, None, 1, out.data_ptr(), None, 0, None)
    assert rc == -1 and b"KS dialect" in lib.pg_last_error()
    rc = lib.pg_fd_lib_gram(U.data_ptr(), 3, 8, 8, 1.0, 1.0, 1.0, 0, 0, 1, 1, 1, None, None, 99, out.data_ptr(), None, 0, None)
    assert rc == -1 and b"n_folds" in lib.pg_last_error()


@pytest.mark.parametrize("tag,kw", [
    ("c1", dict(method="pointwise", dictionary="true")),
    ("c1_sweep", dict(method="pointwise", dictionary="true", grid_search=True)),
    ("c2", dict(method="blockwise", dictionary="true")),
    ("c2_rich_sweep", dict(method="blockwise", dictionary="rich", grid_search=True)),
])
def test_reference_configs_fused(K, ks_default_stack, golden_configs, tag, kw):
    """BASELINE configs[0] and [1] end to end on the GPU (field -> K1 -> K3), against the numbers
    the unmodified reference main() printed and its full-precision replay."""
    U, dx, dy, DT = ks_default_stack
    if not tag.startswith("c1"):
        U = O.add_noise(U, 0.05, seed=999)
    out = K.fit_from_field(U, dx, dy, DT, **kw)
    gold, full = golden_configs[tag], golden_configs["full_precision"][tag]
    assert list(out["X_shape"]) == gold["X_shape"]
    hyper = gold["hyper"]
    assert (out["alpha"], out["threshold"], out["n_active"]) == (hyper["alpha"], hyper["threshold"], hyper["n_active"])
    best_row = [r for r in full["table"] if (r["alpha"], r["threshold"]) == (out["alpha"], out["threshold"])][0]
    assert_coef_close(out["coeffs"], np.array(best_row["coeffs"]), what=tag)
    if tag.startswith("c1"):
        # exact fit (rmse 2e-11 on |y| ~ 0.4): the statistics-derived residual cancels, K3 says so (relres) and the
        # held-out metrics come from the rows (pg_rows_residual_ss).  The reference's sweep then ties on r2 == 1.0
        # and n_active and picks the smallest rmse (ks2d:1731-1741): same cell, same numbers.
        assert out["metrics_source"] == "residuals" and out["metrics_reliable"] and out["relres_min"] < 1e-9
        assert out["r2_test"] == hyper["r2_test"] == 1.0
        np.testing.assert_allclose(out["rmse_test"], hyper["rmse_test"], rtol=1e-4)
    else:
        assert out["metrics_source"] == "statistics"
        np.testing.assert_allclose(out["r2_test"], hyper["r2_test"], rtol=1e-7)
        np.testing.assert_allclose(out["rmse_test"], hyper["rmse_test"], rtol=1e-7)
    if "table" in out and kw.get("grid_search"):
        for (a, thr, r2, err, na), row in zip(out["table"], full["table"]):
            assert (a, thr, na) == (row["alpha"], row["threshold"], row["n_active"])
            np.testing.assert_allclose(r2, row["r2_test"], rtol=1e-7)
            if tag.startswith("c1"):
                np.testing.assert_allclose(err, row["rmse_test"], rtol=1e-4)
            assert_coef_close(out["coef_grid"][O.GRID_ALPHAS.index(a), O.GRID_THRESHOLDS.index(thr)],
                              np.array(row["coeffs"]), what=f"{tag} sweep {a} {thr}")


def test_exact_residuals_from_the_field_on_a_clean_blockwise_fit(K, ks_default_stack):
    """Clean data, blockwise rows (never materialised): the held-out residuals come from a second pass over the field
    (pg_fd_residual_ss) and agree with the oracle's row-based metrics; the selection equals the oracle's."""
    U, dx, dy, DT = ks_default_stack
    U = U[:400]
    out = K.fit_from_field(U, dx, dy, DT, method="blockwise", dictionary="true", grid_search=True)
    ref = O.run_config(U, dx, dy, DT, method="blockwise", dictionary="true", grid_search=True)
    assert out["metrics_source"] == "residuals"
    assert (out["alpha"], out["threshold"]) == (ref["alpha"], ref["threshold"])
    assert_coef_close(out["coeffs"], ref["coeffs"], what="clean blockwise")
    np.testing.assert_allclose(out["rmse_test"], ref["rmse_test"], rtol=1e-3)
    # the residual kernels against explicit rows
    from pde_b200 import _lib as L
    from pde_b200 import ops

    names, X, y = ks_rows(U, dx, dy, DT, "true", False, (3, 8, 8))
    C = np.array([ref["coeffs"], [-1.0, -1.0, -0.5], [0.0, 0.0, 0.0]])
    fold = (np.arange(len(y)) % 3 == 0).astype(np.uint8)
    want = np.array([np.sum((y[fold == 1] - X[fold == 1] @ c) ** 2) for c in C])
    got_rows, n_rows = ops.rows_residual_ss(X, y, C, fold_of_row=fold, eval_fold=1)
    got_fd, n_fd = ops.fd_residual_ss(U, dx, dy, DT, C, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8),
                                      fold_of_row=fold, n_folds=2, eval_fold=1)
    assert n_rows == n_fd == int(fold.sum())
    # residuals of the exact fits are ~3e-12 on |y| ~ 0.4: each carries ~1e-5 relative rounding noise of its own, so
    # their sums agree to ~1e-7 at best; the zero model's sum is an ordinary sum of squares
    # (the second model, the exact PDE, leaves residuals of ~1e-16 each: pure rounding, nothing to compare)
    np.testing.assert_allclose(got_rows[0], want[0], rtol=1e-5)
    np.testing.assert_allclose(got_fd[0], want[0], rtol=1e-3)
    assert got_rows[1] < 1e-25 and got_fd[1] < 1e-25
    np.testing.assert_allclose(got_rows[2], want[2], rtol=1e-12)
    np.testing.assert_allclose(got_fd[2], want[2], rtol=1e-12)


@pytest.mark.parametrize("tag,kw", [
    ("c2_denoise_features", dict(method="blockwise", denoise_time_window=5, denoise_space_sigma=1.5)),
    ("c2_denoise_all_pointwise", dict(method="pointwise", denoise_time_window=3, denoise_space_sigma=3.0, denoise_space_on="all")),
])
def test_fused_path_with_the_denoising_prologue(K, ks_default_stack, golden_configs, tag, kw):
    """ks2d:1448-1468 before the fused path: time moving average, then the periodic Gaussian on the library's stack only
    (u_t from the other stack: pg_fd_lib_gram_two) or on both; against what the reference main() printed."""
    U, dx, dy, DT = ks_default_stack
    U = O.add_noise(U, 0.05, seed=999)
    out = K.fit_from_field(U, dx, dy, DT, dictionary="true", **kw)
    gold = golden_configs[tag]
    assert list(out["X_shape"]) == gold["X_shape"]
    for n, c in zip(out["names"], out["coeffs"]):
        assert abs(c - gold["coeffs_printed"][n]) <= 1.5e-6, (n, c, gold["coeffs_printed"][n])
    np.testing.assert_allclose(out["r2_test"], gold["hyper"]["r2_test"], rtol=1e-6)
    np.testing.assert_allclose(out["rmse_test"], gold["hyper"]["rmse_test"], rtol=1e-6)


def test_statistics_only_fit_flags_unreliable_metrics(K, ks_default_stack):
    """fit_from_stats without an evaluator on an exact fit: no silent clamp, the result says the metrics are noise."""
    from pde_b200 import _lib as L
    from pde_b200 import ops

    U, dx, dy, DT = ks_default_stack
    fof = (np.arange(299) >= 210).astype(np.int32)
    st = ops.fd_lib_gram(U[:300], dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8),
                         fold_of_frame=fof, n_folds=2)
    out = K.fit_from_stats(st[0], st[1], K.TRUE_NAMES)
    assert out["metrics_reliable"] is False and out["metrics_source"] == "statistics" and abs(out["r2_test"] - 1) < 1e-9


def test_reference_script_runs_with_rebound_functions(K, ks_default_stack, golden_configs):
    """Drop-in: a stand-in for the reference module's globals, rebound by patch_reference, then
    driven exactly as main() drives them (ks2d:1508-1550, 1638-1718) for config C2."""
    import types

    import pde_b200

    m = types.SimpleNamespace(**{n: getattr(O, n) for n in ("gradients", "laplacian", "build_dictionary",
                                                            "build_dictionary_true", "build_blockwise_dataset",
                                                            "standardize_fit", "ridge_fit", "stridge")})
    done = pde_b200.patch_reference(m, "ks2d")
    assert set(done) == {"gradients", "laplacian", "build_dictionary", "build_dictionary_true",
                         "build_blockwise_dataset", "standardize_fit", "ridge_fit", "stridge"}
    U, dx, dy, DT = ks_default_stack
    U = O.add_noise(U, 0.05, seed=999)[:301]  # 300 row-frames keep the materialised path small
    rng = np.random.default_rng(0)
    Ut = (U[1:] - U[:-1]) / DT
    names, terms = m.build_dictionary_true(U[:-1], dx=dx, dy=dy, deriv="finite", spectral_cutoff=1.0, include_advection=False)
    X, y = m.build_blockwise_dataset(Ut, terms, names, block_t=3, block_x=8, block_y=8)
    tr, te, scale = O.split_and_scale(names, X, y, rng)
    c = m.stridge(X[tr] / scale, y[tr], alpha=1e-6, threshold=1e-10, max_iter=25) / scale
    ref = O.run_config(U, dx, dy, DT, method="blockwise", dictionary="true")
    assert X.shape == ref["X_shape"]
    assert_coef_close(c, ref["coeffs"], what="rebound C2 (300 frames)")


@pytest.mark.parametrize("shape,block", [((9, 64, 96), (1, 1, 1)), ((13, 128, 64), (3, 8, 8)), ((7, 40, 72), (2, 8, 8))])
def test_slab_additivity_and_variants(ops, L, shape, block):
    """Size-independent properties: statistics are additive over time slabs (what the multi-GPU
    path relies on) and the AUTO variant (tiled where available) equals the generic kernel."""
    U = ops.synth_field(*shape, seed=3, noise=0.05)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_RICH, block=block)
    full = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()[0]
    auto = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_AUTO, **kw).cpu().numpy()[0]
    assert_stats_close(auto, full, 9)
    cut = block[0] * max(1, (shape[0] // 2) // block[0])
    a = ops.fd_lib_gram(U[: cut + 1], 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()[0]
    b = ops.fd_lib_gram(U[cut:], 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()[0]
    assert_stats_close(a + b, full, 9)
    # and against the oracle on the same device-generated field
    names, X, y = ks_rows(U.cpu().numpy(), 0.5, 0.5, 1e-3, "rich", False, block)
    assert_stats_close(full, gram.pack_stats(X, y), 9)


def test_stridge_sign_constrained_matches_reference():
    """K3 with sign constraints (ks2d:552-600) against the reference's outputs on the golden rows: identical
    support, coefficients to 1e-8, through the literal (X, y) signature."""
    from conftest import GOLDEN
    from pde_b200 import ks2d as K

    gs, g = np.load(GOLDEN / "ks2d_signed.npz"), np.load(GOLDEN / "ks2d_small.npz")
    for tag, xkey, ykey in [("true_pw", "bw111_X_true", "bw111_y"), ("rich_pw", "bw111_X_rich", "bw111_y"),
                            ("rich_453", "bw453_X_rich", "bw453_y")]:
        X, y = g[xkey], g[ykey]
        for signs, (a, t), ref in zip(gs[f"{tag}_signs"], gs[f"{tag}_grid"], gs[f"{tag}_coef"]):
            got = K.stridge_sign_constrained(X, y, alpha=a, threshold=t, max_iter=25, signs=[int(v) for v in signs])
            assert_coef_close(got, ref, what=f"{tag} signs={signs} alpha={a} thr={t}")
        assert_coef_close(K.stridge_sign_constrained(X, y, alpha=1e-3, threshold=1e-6, signs=None), gs[f"{tag}_none"])
        assert_coef_close(K.stridge_sign_constrained(X, y, alpha=1e-3, threshold=1e-6, max_iter=0, signs=[-1] * X.shape[1]),
                          gs[f"{tag}_iter0"])
    with pytest.raises(ValueError):
        K.stridge_sign_constrained(g["bw111_X_true"], g["bw111_y"], signs=[2, 0, 0])


@pytest.mark.parametrize("tag", ["c1", "c2", "c2_rich_sweep"])
def test_rollout_on_gpu_matches_reference_main(tag, golden_configs, ks_default_stack):
    """pg_ks_rollout (explicit Euler with the fitted right-hand side, ks2d:1804-1838) against the per-step
    RMSEs of the reference's own main() and against the oracle: same arithmetic per point (bit-identical
    u_hat), only the order of the squared-error sum differs."""
    import json

    from conftest import GOLDEN
    from pde_b200 import ks2d as K
    from helpers import rollout_case

    g = json.loads((GOLDEN / "ks2d_rollout.json").read_text())[tag]
    Uo, dx, dy, DT, names, coef = rollout_case(tag, golden_configs, ks_default_stack)
    errs = K.rollout_errors(Uo, dx, dy, DT, names, coef, 50)
    np.testing.assert_allclose(errs, g["errs"], rtol=1e-10, atol=0)
    np.testing.assert_allclose(errs, O.rollout_errors(Uo, dx, dy, DT, names, coef, 50), rtol=1e-10, atol=0)
    assert K.rollout_errors(Uo[:4], dx, dy, DT, names, coef, 50).shape == (3,)      # ks2d:1832: min(steps, T-1)
    with pytest.raises(ValueError):
        K.rollout_errors(Uo[:4], dx, dy, DT, ["u", "nonsense"], [1.0, 2.0])


@pytest.mark.parametrize("split", ["left_right", "top_bottom"])
def test_spatial_holdout_folds(split):
    """Spatial hold-out regions (analyze_results:282-299) through the tiled block kernel's per-row folds: the
    statistics of each region equal the oracle's rows restricted to that region, and the fit follows."""
    from pde_b200 import _lib as L
    from pde_b200 import ks2d as K
    from pde_b200 import ops

    U = ops.synth_field(13, 128, 256, seed=31, noise=0.05)
    Uh = U.cpu().numpy()
    block = (3, 8, 8)
    fold = K.spatial_fold_of_row(12, 128, 256, block, split=split, train_frac=0.7)
    stats = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=block,
                            fold_of_row=fold, n_folds=2, variant=L.VARIANT_TILED).cpu().numpy()
    _, X, y = ks_rows(Uh, 0.5, 0.5, 1e-3, "true", False, block)
    jb, ib = np.meshgrid(np.arange(32), np.arange(16))
    held = (jb * 8 >= K.split_space(256, 0.7)) if split == "left_right" else (ib * 8 >= K.split_space(128, 0.7))
    held = np.broadcast_to(held[None], (4, 16, 32)).reshape(-1)
    assert np.array_equal(held.astype(np.uint8), fold)
    for f in range(2):
        assert_stats_close(stats[f], gram.pack_stats(X[held == f], y[held == f]), 3)
    out = K.fit_from_field(U, 0.5, 0.5, 1e-3, method=f"blockwise_{split}", dictionary="true", grid_search=True)
    ref = gram.ks_fit_from_stats(gram.pack_stats(X[~held], y[~held]), gram.pack_stats(X[held], y[held]), 3,
                                 alphas=K.GRID_ALPHAS, thresholds=K.GRID_THRESHOLDS)
    assert (out["alpha"], out["threshold"]) == (ref["alpha"], ref["threshold"])
    assert_coef_close(out["coeffs"], ref["coeffs"], what=split)
    with pytest.raises(ValueError):
        K.split_space(100, 0.95)


def test_smoothing_prologue_on_gpu():
    """pg_time_moving_average is bit-identical to the reference's cumulative-sum formulation; the two-pass
    periodic stencil reproduces gaussian_smooth_periodic_2d (FFT) to rounding (ks2d:125-161)."""
    from conftest import GOLDEN
    from pde_b200 import ks2d as K

    g = np.load(GOLDEN / "ks2d_smooth.npz")
    for w in (3, 5, 9):
        assert np.array_equal(K.time_smooth_moving_average(g["stack"], w), g[f"tavg_{w}"])
    assert np.array_equal(K.time_smooth_moving_average(g["stack"], 1), g["stack"])
    with pytest.raises(ValueError):
        K.time_smooth_moving_average(g["stack"], 4)
    for tag in "abc":
        f = g[f"frame_{tag}"]
        for sig in (0.8, 1.5, 4.0):
            ref = g[f"gauss_{tag}_{sig}"]
            got = K.gaussian_smooth_periodic_2d(f, sig)
            assert np.abs(got - ref).max() <= 1e-13 * np.abs(ref).max(), (tag, sig)
    assert np.array_equal(K.gaussian_smooth_periodic_2d(g["frame_a"], 0.0), g["frame_a"])


def test_time_moving_average_register_ring_and_plain_kernels():
    """Windows 3 / 5 / 7 / 9 on frames with an even number of points take the register-ring kernel (one read of the
    stack), everything else (window 11, odd frames, T < window) the plain one: both bit-identical to the oracle's
    cumulative-sum formulation (ks2d:145-161)."""
    from pde_b200 import ks2d as K

    rng = np.random.default_rng(5)
    for T, A0, A1 in ((23, 6, 10), (23, 5, 7), (4, 6, 10), (9, 8, 8), (10, 3, 4)):
        U = rng.standard_normal((T, A0, A1))
        for w in (3, 5, 7, 9, 11):
            if w // 2 > T - 1:          # np.pad(mode="reflect") needs pad <= T - 1
                continue
            assert np.array_equal(K.time_smooth_moving_average(U, w), O.time_smooth_moving_average(U, w)), (T, A0, A1, w)


def test_time_holdout_cross_validation():
    """K = 5 time-holdout folds from one K1 pass; K x 30 fits in one K3 launch; every fold's fit and held-out
    score against the oracle on the rows of that fold ("also K = 5")."""
    from pde_b200 import ks2d as K
    from pde_b200 import ops

    U = ops.synth_field(31, 64, 128, seed=41, noise=0.05)
    Uh = U.cpu().numpy()
    block = (3, 8, 8)
    out = K.fit_time_cv(U, 0.5, 0.5, 1e-3, n_folds=5, dictionary="true", block=block)
    fof = out["fold_of_frame"]
    assert fof.shape == (30,) and set(fof) == {0, 1, 2, 3, 4} and all(len(set(fof[3 * k:3 * k + 3])) == 1 for k in range(10))
    _, X, y = ks_rows(Uh, 0.5, 0.5, 1e-3, "true", False, block)
    fold_of_row = np.repeat(fof[::3], 8 * 16)
    for k in range(5):
        tr, te = fold_of_row != k, fold_of_row == k
        assert_stats_close(out["stats"][k], gram.pack_stats(X[te], y[te]), 3)
        ref = gram.ks_fit_from_stats(gram.pack_stats(X[tr], y[tr]), gram.pack_stats(X[te], y[te]), 3,
                                     alphas=K.GRID_ALPHAS, thresholds=K.GRID_THRESHOLDS)
        ia, it = out["best_per_fold"][k]
        assert (K.GRID_ALPHAS[ia], K.GRID_THRESHOLDS[it]) == (ref["alpha"], ref["threshold"])
        assert_coef_close(out["coef_grid"][k, ia, it], ref["coeffs"], what=f"fold {k}")
        np.testing.assert_allclose(out["metrics"][k, ia, it, 0], ref["r2_test"], rtol=1e-6, atol=1e-9)


def test_ensemble_stridge_weighted_grams(golden_ks2d):
    """ensemble_stridge (ks2d:603-642): every bootstrap resample is fitted from the multiplicity-weighted Gram of
    the original rows; per-resample coefficients against the oracle's fits of the materialised resamples, and
    (median, std) against the reference's outputs."""
    from conftest import GOLDEN
    from pde_b200 import _lib as L
    from pde_b200 import ks2d as K
    from pde_b200 import ops

    g = np.load(GOLDEN / "ks2d_ensemble.npz")
    y = golden_ks2d["bw111_y"]
    for tag, X in (("true", golden_ks2d["bw111_X_true"]), ("rich", golden_ks2d["bw111_X_rich"])):
        for k in range(2):
            a, t, nb, frac, seed = g[f"{tag}_{k}_args"]
            kw = dict(alpha=a, threshold=t, n_bootstrap=int(nb), subsample_frac=frac, seed=int(seed))
            med, std = K.ensemble_stridge(X, y, **kw)
            assert_coef_close(med, g[f"{tag}_{k}_median"], what=f"ensemble median {tag} {k}")
            np.testing.assert_allclose(std, g[f"{tag}_{k}_std"], rtol=1e-6, atol=1e-12)
    # weighted statistics == statistics of the materialised resample
    X = golden_ks2d["bw111_X_rich"]
    rng = np.random.default_rng(5)
    idx = rng.choice(len(y), size=2000, replace=True)
    w = np.bincount(idx, minlength=len(y))[None]
    st = ops.rows_gram_weighted(X, y, w).cpu().numpy()[0]
    assert_stats_close(st, gram.pack_stats(X[idx], y[idx]), 9)
    with pytest.raises(NotImplementedError):
        K.ensemble_stridge(X, y, use_huber=True)


def test_staged_host_device_copies_round_trip():
    """Large NumPy arrays travel through pinned staging buffers (pde_b200._xfer): bit-identical both ways, for sizes
    around the chunk boundaries, non-contiguous and read-only inputs and other dtypes."""
    import torch

    from pde_b200 import _xfer

    rng = np.random.default_rng(0)
    per = _xfer.STAGE_BYTES // 8
    for n in (1000, _xfer.MIN_BYTES // 8, per - 1, per, per + 1, 2 * per + 12345, 5 * per // 2):
        a = rng.standard_normal(n)
        d = _xfer.to_device(a)
        assert d.is_cuda and d.dtype == torch.float64 and tuple(d.shape) == a.shape
        assert np.array_equal(d.cpu().numpy(), a)
        assert np.array_equal(_xfer.to_host(d * 1.0), a)
    a = rng.standard_normal((300, 200, 120)).astype(np.float32)      # 28.8 MB, fp32, 3-D
    a.flags.writeable = False
    assert np.array_equal(_xfer.to_host(_xfer.to_device(a)), a)
    b = rng.standard_normal((64, 512, 512))[:, ::2, :]               # non-contiguous view, 67 MB
    got = _xfer.to_host(_xfer.to_device(b))
    assert got.shape == b.shape and np.array_equal(got, b)
    i = rng.integers(0, 2 ** 40, size=3_000_000)                     # int64
    assert np.array_equal(_xfer.to_host(_xfer.to_device(i)), i)


def test_fit_streamed_from_host_stacks():
    """slabs.fit_streamed: a HOST stack streamed in slabs (pinned tensor: DMA straight from it; NumPy array: pinned
    staging) gives the statistics of the device-resident stack (slab additivity), folds included."""
    import torch

    from pde_b200 import _lib as L
    from pde_b200 import ops, slabs

    U = ops.synth_field(200, 256, 256, seed=4, noise=0.05)
    fof = (np.arange(199) >= 138).astype(np.int32)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8), n_folds=2)
    ref = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_frame=fof, **kw).cpu().numpy()
    host_np = U.cpu().numpy()
    pinned = torch.from_numpy(host_np).pin_memory()
    for src in (host_np, pinned, torch.from_numpy(host_np)):
        got = slabs.fit_streamed(src, 0.5, 0.5, 1e-3, fold_of_frame=fof, slab_frames=48, **kw).cpu().numpy()
        for f in range(2):
            assert_stats_close(got[f], ref[f], 3, rtol=1e-12)


@pytest.mark.parametrize("shape,sigma", [((3, 128, 192), 1.0), ((2, 256, 130), 0.8), ((2, 144, 128), 3.5), ((1, 129, 131), 1.5)])
def test_periodic_gaussian_fft_route(shape, sigma):
    """pg_periodic_gaussian_fft (the reference's own formulation of gaussian_smooth_periodic_2d, ks2d:125-142: an FFT
    product) against the oracle's NumPy FFT and against the direct circular convolution (pg_periodic_conv), on frames
    with even and odd extents; 1e-13 of the frame's magnitude like the direct route."""
    from oracle import ks2d as O
    from pde_b200 import ops

    rng = np.random.default_rng(shape[1])
    U = rng.standard_normal(shape) + 0.3
    fft = ops.gaussian_smooth_periodic(U, sigma, method="fft").cpu().numpy()
    conv = ops.gaussian_smooth_periodic(U, sigma, method="conv").cpu().numpy()
    auto = ops.gaussian_smooth_periodic(U, sigma).cpu().numpy()
    for t in range(shape[0]):
        ref = O.gaussian_smooth_periodic_2d(U[t], sigma)
        scale = np.abs(ref).max()
        assert np.abs(fft[t] - ref).max() <= 1e-13 * scale
        assert np.abs(conv[t] - ref).max() <= 1e-13 * scale
        assert np.abs(auto[t] - ref).max() <= 1e-13 * scale
