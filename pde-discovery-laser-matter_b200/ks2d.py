"""Drop-in for the hot-path functions of scripts/ks2d_stridge_benchmark.py ("ks2d").

Same names, argument meaning and error behaviour as the reference; every function runs on
the GPU through libpdegram.so and returns NumPy arrays of the reference's shape and dtype.
``fit_from_field`` is the fused fast path (field -> statistics -> STRidge, Theta never
materialised) that restates the reference's main() hot path (ks2d:1508-1743).

Layout: ``U[t, x, y]`` with "x" on frame axis 0 (ks2d:70-73, ks2d:1295).
"""

from __future__ import annotations

import numpy as np

from . import _lib as L
from . import ops

RICH_NAMES = ["1", "u", "u^2", "u_x", "u_y", "∇²u", "∇⁴u", "|∇u|²", "u·∇²u"]  # ks2d:1048-1059
TRUE_NAMES = ["∇²u", "∇⁴u", "|∇u|²"]                                         # ks2d:1095-1099
ADV_NAMES = TRUE_NAMES + ["u_x", "u_y"]                                       # ks2d:1100-1102
RICH_NOADV_NAMES = [n for n in RICH_NAMES if n not in ("u_x", "u_y")]         # ks2d:1536-1539
GRID_ALPHAS = (1e-6, 1e-5, 1e-4, 1e-3, 1e-2)                                  # ks2d:1721
GRID_THRESHOLDS = (1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5)                       # ks2d:1722

_LIB_OF = {"true": (L.LIB_KS_TRUE, TRUE_NAMES), "true_adv": (L.LIB_KS_TRUE_ADV, ADV_NAMES),
           "rich": (L.LIB_KS_RICH, RICH_NAMES), "rich_noadv": (L.LIB_KS_RICH_NOADV, RICH_NOADV_NAMES)}


def _np(t):
    from . import _xfer

    return _xfer.to_host(t)       # large results: pinned double-buffered staging


def _frame(f2d):
    f = np.asarray(f2d, dtype=np.float64)
    if f.ndim != 2:
        raise ValueError("expected a 2-D frame")
    return f[None]


# ------------------------------------------------------------------ fit metrics (ks2d:29-40)
def rmse(y_true, y_pred) -> float:
    """ks2d:29-32."""
    s, n = ops.fit_metric_sums(y_true, y_pred)
    return float(np.sqrt(s[1] / n)) if n else float("nan")


def r2_score(y_true, y_pred) -> float:
    """ks2d:35-40 (with its 1e-18 guard)."""
    s, n = ops.fit_metric_sums(y_true, y_pred)
    return float(1.0 - s[1] / (s[5] + 1e-18)) if n else float("nan")


def standardize_transform(X, mean, scale):
    """ks2d:51-52: (X - mean) / scale (elementwise; NumPy broadcasting on the host, nothing to accelerate)."""
    return (np.asarray(X) - np.asarray(mean)) / np.asarray(scale)


# ------------------------------------------------------------------ stencils (ks2d:63-73)
def laplacian(f2d, dx: float, dy: float):
    """ks2d:63-67: periodic 5-point Laplacian of one frame."""
    return _np(ops.fd_terms(_frame(f2d), dx, dy, 1.0, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_LAP))[0, 0]


def gradients(f2d, dx: float, dy: float):
    """ks2d:70-73: periodic central differences (gx along axis 0, gy along axis 1)."""
    g = _np(ops.fd_terms(_frame(f2d), dx, dy, 1.0, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_GRAD))
    return g[0, 0], g[1, 0]


# ------------------------------------------------------------------ optional denoising prologue (ks2d:125-161)
def gaussian_smooth_periodic_2d(frame, sigma_px: float):
    """ks2d:125-142: periodic Gaussian low-pass of one frame (sigma in pixels)."""
    f = np.asarray(frame)
    if float(sigma_px) <= 0:
        return f.astype(np.float64, copy=True)
    return _np(ops.gaussian_smooth_periodic(_frame(f), sigma_px))[0]


def time_smooth_moving_average(U, window: int):
    """ks2d:145-161: reflect-padded moving average along axis 0; window must be odd."""
    window = int(window)
    U = np.asarray(U)
    if window <= 1:
        return U.astype(np.float64, copy=True)
    if window % 2 == 0:
        raise ValueError("time smoothing window must be odd")
    return _np(ops.time_moving_average(U, window))


# ------------------------------------------------------------------ dictionaries (ks2d:1017-1104)
def _dictionary(U, dx, dy, deriv, key):
    if deriv != "finite":
        raise NotImplementedError("only deriv='finite' is GPU-accelerated; the spectral path has no CPU fallback here")
    lib, names = _LIB_OF[key]
    terms = _np(ops.fd_terms(U, dx, dy, 1.0, dialect=L.FD_KS_PERIODIC, library=lib))
    return list(names), {n: terms[k] for k, n in enumerate(names)}


def build_dictionary(U_mid, dx: float, dy: float, *, deriv: str = "finite", spectral_cutoff: float = 1.0):
    """ks2d:1017-1060: the p=9 rich dictionary; returns (names, {name: (T,Nx,Ny) array})."""
    return _dictionary(U_mid, dx, dy, deriv, "rich")


def build_dictionary_true(U_frames, dx: float, dy: float, *, deriv: str = "finite", spectral_cutoff: float = 1.0,
                          include_advection: bool = False):
    """ks2d:1063-1104: [lap, bih, |grad|^2] (+ u_x, u_y)."""
    return _dictionary(U_frames, dx, dy, deriv, "true_adv" if include_advection else "true")


# ------------------------------------------------------------------ blockwise means (ks2d:358-401)
def build_blockwise_dataset(Ut, terms, names, *, block_t: int, block_x: int, block_y: int):
    """ks2d:358-401: mean of Ut and of each term over (bt,bx,by) blocks, ragged trailing
    blocks kept, rows with a non-finite mean dropped."""
    Ut = np.asarray(Ut)
    if Ut.ndim != 3:
        raise ValueError("Ut must be (T, Nx, Ny)")
    bt, bx, by = int(block_t), int(block_x), int(block_y)
    if bt <= 0 or bx <= 0 or by <= 0:
        raise ValueError("block sizes must be > 0")
    torch = L.torch_cuda()
    from . import _xfer

    # [1 + p][T][Nx][Ny] device buffer filled array by array (copies only; pg_block_means takes the stack)
    stack = torch.empty((1 + len(names),) + Ut.shape, dtype=torch.float64, device="cuda")
    for k, a in enumerate([Ut] + [terms[n] for n in names]):
        a = np.asarray(a, dtype=np.float64)
        if a.shape != Ut.shape:
            raise ValueError("every term must have the shape of Ut")
        _xfer.to_device(a, out=stack[k])
    rows = _np(ops.block_means(stack, (bt, bx, by)))
    keep = np.isfinite(rows).all(axis=1)
    if not keep.any():
        return np.zeros((0, len(names)), dtype=np.float64), np.zeros((0,), dtype=np.float64)
    return rows[keep, 1:].copy(), rows[keep, 0].copy()


# ------------------------------------------------------------------ STRidge (ks2d:43-60, 404-428)
def _stats_of_rows(X, y):
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if X.ndim != 2 or y.shape != (X.shape[0],):
        raise ValueError("X must be (n, p) and y (n,)")
    if X.shape[1] > L.PG_MAX_P:
        raise ValueError(f"at most {L.PG_MAX_P} columns are supported")
    shift = X[:1].copy()
    stats, mm = ops.rows_gram(X, y, shift=shift, want_minmax=True)
    return stats[:, 0], mm[:, 0], shift


def standardize_fit(X):
    """ks2d:43-48: (mean, population std with 0 -> 1) per column, from GPU statistics."""
    X = np.asarray(X, dtype=np.float64)
    stats, mm, shift = _stats_of_rows(X, np.zeros(X.shape[0]))
    s, mm = _np(stats)[0], _np(mm)[0]
    p, n = X.shape[1], s[0]
    mu = s[3:3 + p] / n
    diag = np.array([s[3 + 2 * p + i * p - (i * (i - 1)) // 2] for i in range(p)])
    var = np.maximum(diag / n - mu * mu, 0.0)
    scale = np.sqrt(var)
    scale[(mm[0] == mm[1]) | ~(scale > 0)] = 1.0
    return mu + shift[0], scale


def ridge_fit(X, y, alpha: float):
    """ks2d:55-60: solve (X^T X + alpha I) b = X^T y (LU with partial pivoting on the GPU)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    stats = ops.rows_gram(X, np.asarray(y, dtype=np.float64))[:, 0]
    out = ops.stridge_batched(stats, X.shape[1], dialect=L.STRIDGE_BASIC, alphas=[alpha], thresholds=[0.0], max_iter=1)
    return _np(out["coef"])[0, 0, 0]


def stridge(X, y, *, alpha: float = 1e-3, threshold: float = 1e-6, max_iter: int = 25):
    """ks2d:404-428 on the GPU: rows -> statistics (pg_rows_gram) -> K3 (pg_stridge_batched)."""
    stats, mm, shift = _stats_of_rows(X, y)
    out = ops.stridge_batched(stats, np.asarray(X).shape[1], dialect=L.STRIDGE_KS, alphas=[alpha],
                              thresholds=[threshold], max_iter=int(max_iter), colminmax=mm, shift=shift)
    return _np(out["coef"])[0, 0, 0]


def stridge_sign_constrained(X, y, *, alpha: float = 1e-3, threshold: float = 1e-6, max_iter: int = 25, signs=None):
    """ks2d:552-600 on the GPU: STRidge with physics-informed sign constraints (signs[j] in {-1, 0, +1};
    None = unconstrained).  Rows -> statistics -> K3 with the sign filter inside the thresholding loop."""
    stats, mm, shift = _stats_of_rows(X, y)
    out = ops.stridge_batched(stats, np.asarray(X).shape[1], dialect=L.STRIDGE_KS, alphas=[alpha],
                              thresholds=[threshold], max_iter=int(max_iter), colminmax=mm, shift=shift,
                              signs=None if signs is None else list(signs))
    return _np(out["coef"])[0, 0, 0]


def ensemble_stridge(X, y, *, alpha: float = 1e-3, threshold: float = 1e-6, max_iter: int = 25, n_bootstrap: int = 50,
                     subsample_frac: float = 0.7, seed: int = 0, use_huber: bool = False, huber_delta: float = 1.35):
    """ks2d:603-642 on the GPU: the reference draws ``rng.choice(n, n_sub, replace=True)`` per bootstrap and
    fits stridge() on the materialised resample; a resample's Gram is the multiplicity-weighted Gram of the
    original rows, so here the draws stay on the host RNG (same stream, same order), become a (n_bootstrap, n)
    count table, one launch forms all weighted statistics and one K3 launch fits all resamples.  Returns
    (median, std) over the resamples like the reference.  ``use_huber`` needs raw residuals: not supported."""
    if use_huber:
        raise NotImplementedError("the Huber inner solve iterates on raw residuals; only the ridge ensemble is GPU-accelerated")
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if X.ndim != 2 or y.shape != (X.shape[0],):
        raise ValueError("X must be (n, p) and y (n,)")
    rng = np.random.default_rng(seed)
    n, p = X.shape
    n_sub = max(int(n * subsample_frac), 1)
    counts = np.stack([np.bincount(rng.choice(n, size=n_sub, replace=True), minlength=n) for _ in range(int(n_bootstrap))])
    shift = X[:1].copy()
    stats, mm = ops.rows_gram_weighted(X, y, counts.astype(np.uint16), shift=shift, want_minmax=True)
    out = ops.stridge_batched(stats, p, dialect=L.STRIDGE_KS, alphas=[alpha], thresholds=[threshold], max_iter=int(max_iter),
                              colminmax=mm, shift=np.ascontiguousarray(np.broadcast_to(shift, (int(n_bootstrap), p))))
    C = _np(out["coef"])[:, 0, 0, :]
    return np.median(C, axis=0), np.std(C, axis=0)


def rollout_errors(U, dx, dy, DT, names, coeffs, n_steps: int = 50):
    """ks2d:1804-1838: explicit-Euler rollout of the discovered PDE from U[0]; RMSE against U[k+1] per step.
    ``names`` selects the library (true / true+advection / rich / rich without advection)."""
    names = list(names)
    lib = next((l for l, nm in _LIB_OF.values() if list(nm) == names), None)
    if lib is None:
        raise ValueError(f"no KS library with columns {names}")
    n = int(min(n_steps, np.asarray(U).shape[0] - 1))       # ks2d:1832
    return _np(ops.ks_rollout(U, dx, dy, DT, coeffs, n, library=lib))


# ------------------------------------------------------------------ fused path (ks2d:1508-1743)
def library_of(dictionary: str, include_advection: bool = False, enforce_no_advection: bool = False):
    if dictionary == "true":
        return _LIB_OF["true_adv" if (include_advection and not enforce_no_advection) else "true"]
    return _LIB_OF["rich_noadv" if enforce_no_advection else "rich"]


def split_folds(n_rows: int, rng):
    """ks2d:1638-1641 as a fold vector: 0 = train (first 70 % of the permutation), 1 = test."""
    perm = rng.permutation(n_rows)
    split = int(0.7 * n_rows)
    fold = np.ones(n_rows, dtype=np.uint8)
    fold[perm[:split]] = 0
    return fold, perm


def split_space(extent: int, train_frac: float) -> int:
    """analyze_results:282-299: index where the spatial hold-out region starts (floor(frac * extent), clipped so
    that both regions are non-empty)."""
    if not (0.4 <= train_frac <= 0.9):
        raise ValueError("SPACE_TRAIN_FRAC should be in [0.4, 0.9]")
    return max(1, min(extent - 1, int(np.floor(train_frac * extent))))


def spatial_fold_of_row(n_row_frames: int, A0: int, A1: int, block=(1, 1, 1), *, split="left_right", train_frac=0.7):
    """Spatial hold-out folds (analyze_results:282-299, 820-902) as one fold byte per (block) row in K1's row
    order: 0 = train region, 1 = held-out region.  "left_right" splits the LAST frame axis (columns),
    "top_bottom" the first one (rows); a block belongs to the region its first point lies in."""
    bt, b0, b1 = (int(b) for b in block)
    nbt, nb0, nb1 = -(-n_row_frames // bt), -(-A0 // b0), -(-A1 // b1)
    if split == "left_right":
        held = (np.arange(nb1) * b1 >= split_space(A1, train_frac))[None, None, :]
    elif split == "top_bottom":
        held = (np.arange(nb0) * b0 >= split_space(A0, train_frac))[None, :, None]
    else:
        raise ValueError("split must be 'left_right' or 'top_bottom'")
    return np.broadcast_to(held, (nbt, nb0, nb1)).astype(np.uint8).reshape(-1)


# below this ratio ss_res / sum y^2 the statistics-derived residual (yy - 2 c.b + c.G.c) is rounding noise
CANCELLATION_RELRES = 1e-9


def fit_from_stats(stats_train, stats_test, names, *, alpha=1e-6, threshold=1e-10, grid_search=False, max_iter=25,
                   signs=None, exact_residuals=None):
    """ks2d:1647-1779 on statistics: train-RMS scale, STRidge (or the 5x6 sweep), held-out
    r2/rmse and the reference's arg-max, all inside pg_stridge_batched.

    Held-out metrics from statistics cancel when a fit is exact to ~1e-8 (the reference's clean configs): K3 reports
    the ratio ss_res / sum y^2, and when any cell falls below ``CANCELLATION_RELRES`` the residual sums are taken from
    the rows instead -- ``exact_residuals(coef [J][p]) -> (ss_res [J], n_rows)`` (``ops.rows_residual_ss`` /
    ``ops.fd_residual_ss`` bound to the data) -- and r2, rmse and the arg-max (key (r2, -n_active, -rmse), first
    maximum wins: ks2d:1731-1741) are redone from them.  Without an evaluator the result says
    ``metrics_reliable=False`` instead of passing noise off as an rmse."""
    p = len(names)
    const_cols = [j for j, n in enumerate(names) if n == "1"]
    alphas = GRID_ALPHAS if grid_search else (alpha,)
    thrs = GRID_THRESHOLDS if grid_search else (threshold,)
    out = ops.stridge_batched(stats_train, p, dialect=L.STRIDGE_KS, flags=L.STRIDGE_RMS_PRESCALE, alphas=alphas,
                              thresholds=thrs, max_iter=max_iter, const_cols=const_cols, eval_stats=stats_test,
                              signs=signs)
    coef, met, best = _np(out["coef"])[0], _np(out["metrics"])[0].copy(), int(_np(out["best"])[0])
    relres = _np(out["relres"])[0]
    reliable, source = True, "statistics"
    if float(np.min(relres)) < CANCELLATION_RELRES:
        if exact_residuals is None:
            reliable = False
        else:
            st = _np(ops._dev(stats_test)).reshape(-1)
            n_te, sy, syy = st[0], st[1], st[2]
            ss, n_rows = exact_residuals(coef.reshape(-1, p))
            if n_rows != int(n_te):
                raise L.PdeGramError(f"the residual pass saw {n_rows} held-out rows, the statistics {int(n_te)}")
            ss_tot = syy - sy * sy / n_te
            met[..., 0] = (1.0 - ss / (ss_tot + 1e-18)).reshape(len(alphas), len(thrs))      # ks2d:35-40
            met[..., 1] = np.sqrt(ss / n_te).reshape(len(alphas), len(thrs))                  # ks2d:29-32
            key_best, best = None, 0
            for k in range(len(alphas) * len(thrs)):                                          # ks2d:1731-1741
                a, t = divmod(k, len(thrs))
                key = (met[a, t, 0], -int(np.sum(np.abs(coef[a, t]) > 0)), -met[a, t, 1])
                if key_best is None or key > key_best:
                    key_best, best = key, k
            source = "residuals"
    ia, it = divmod(best, len(thrs))
    c = coef[ia, it]
    return dict(names=list(names), alpha=alphas[ia], threshold=thrs[it], coeffs=c, r2_test=float(met[ia, it, 0]),
                rmse_test=float(met[ia, it, 1]), n_active=int(np.sum(np.abs(c) > 0)),
                table=[(alphas[a], thrs[t], float(met[a, t, 0]), float(met[a, t, 1]),
                        int(np.sum(np.abs(coef[a, t]) > 0))) for a in range(len(alphas)) for t in range(len(thrs))],
                coef_grid=coef, metrics_reliable=reliable, metrics_source=source, relres_min=float(np.min(relres)))


def time_fold_of_frame(n_row_frames: int, n_folds: int, block_t: int = 1):
    """K contiguous time-holdout folds: fold id per row frame, boundaries aligned to whole
    t-blocks so that no block straddles two folds."""
    n_tb = -(-n_row_frames // int(block_t))
    tb_fold = np.minimum(np.arange(n_tb) * int(n_folds) // max(n_tb, 1), n_folds - 1)
    return np.repeat(tb_fold, int(block_t))[:n_row_frames].astype(np.int32)


def fit_time_cv(U, dx, dy, DT, *, n_folds=5, dictionary="true", include_advection=False, enforce_no_advection=False,
                block=(3, 8, 8), max_iter=25, variant=L.VARIANT_AUTO):
    """K-fold time-holdout cross-validation of the 5 x 6 (alpha, threshold) sweep (ks2d:1720-1743 applied per
    fold): ONE pass of K1 over the stack yields the statistics of the K time folds (they cost nothing extra),
    the K train sets are sums of the other folds' statistics, and one K3 launch fits K x 30 models and scores
    each on its held-out fold.  Returns the per-fold selections and the configuration with the best mean r2."""
    lib, names = library_of(dictionary, include_advection, enforce_no_advection)
    torch = L.torch_cuda()
    Ud = ops.field(U)
    T = Ud.shape[0]
    fof = time_fold_of_frame(T - 1, n_folds, block[0])
    stats = ops.fd_lib_gram(Ud, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=lib, block=block, fold_of_frame=fof,
                            n_folds=n_folds, variant=variant)                      # [K][S]
    p = len(names)
    train = torch.zeros_like(stats)                     # train set of fold k = sum of the other folds' statistics
    for k_ in range(n_folds):
        for j in range(n_folds):
            if j != k_:
                ops.stats_accumulate(train[k_], stats[j])
    const_cols = [j for j, n in enumerate(names) if n == "1"]
    out = ops.stridge_batched(train, p, dialect=L.STRIDGE_KS, flags=L.STRIDGE_RMS_PRESCALE, alphas=GRID_ALPHAS,
                              thresholds=GRID_THRESHOLDS, max_iter=max_iter, const_cols=const_cols, eval_stats=stats)
    coef, met, best = _np(out["coef"]), _np(out["metrics"]), _np(out["best"])
    mean_r2 = met[..., 0].mean(axis=0)                                             # [na][nt]
    ia, it = np.unravel_index(int(np.argmax(mean_r2)), mean_r2.shape)
    return dict(names=list(names), fold_of_frame=fof, stats=_np(stats), coef_grid=coef, metrics=met,
                best_per_fold=[divmod(int(b), len(GRID_THRESHOLDS)) for b in best], mean_r2=mean_r2,
                alpha=GRID_ALPHAS[ia], threshold=GRID_THRESHOLDS[it], coeffs=coef[:, ia, it].mean(axis=0))


def fit_from_field(U, dx, dy, DT, *, method="blockwise", dictionary="true", include_advection=False,
                   enforce_no_advection=False, block=(3, 8, 8), n_sample=50_000, alpha=1e-6, threshold=1e-10,
                   grid_search=False, fold_of_frame=None, seed=0, variant=L.VARIANT_AUTO, signs=None,
                   denoise_time_window=1, denoise_space_sigma=0.0, denoise_space_on="features"):
    """The reference's main() hot path for one config, fused on the GPU.

    ``denoise_*``: the script's optional smoothing before the path (ks2d:1448-1468): a reflect-padded moving average
    along t, then a periodic Gaussian per frame applied to the stack the library is built from ("features": u_t still
    comes from the time-smoothed stack, so K1 reads two stacks, generic kernel) or to both ("all": one smoothed stack,
    the tiled kernels).  The smoothing runs as its own kernels before K1 (a separate pass over HBM; see DESIGN.md on
    why it is not fused into K1's shared-memory tiles).

    method="blockwise": K1 forms block means and both folds' Grams in one pass over U; the
        70/30 row permutation (ks2d:1638-1641) stays on the host RNG and is passed as one fold
        byte per block row.
    method="pointwise": the 50 000-point sample (ks2d:1625-1636) is gathered by K1c, then split.
    method="full": every grid point is a row and ``fold_of_frame`` gives time-holdout folds
        (0 = train, 1 = test); this is the large-stack configuration C4/C5.
    method="blockwise_left_right" / "blockwise_top_bottom": block means with a SPATIAL hold-out region
        (analyze_results:282-299, 820-902) instead of the random row split.
    ``signs``: optional sign constraints of stridge_sign_constrained (ks2d:552-600).
    """
    lib, names = library_of(dictionary, include_advection, enforce_no_advection)
    torch = L.torch_cuda()
    Ud = ops.field(U)
    T, A0, A1 = Ud.shape
    Uy = None                                   # the stack u_t is taken of, when it differs from the library's
    if int(denoise_time_window) > 1:
        if int(denoise_time_window) % 2 == 0:
            raise ValueError("time smoothing window must be odd")
        Ud = ops.time_moving_average(Ud, int(denoise_time_window))
    if float(denoise_space_sigma) > 0.0:
        if denoise_space_on == "all":
            Ud = ops.gaussian_smooth_periodic(Ud, float(denoise_space_sigma))
        elif denoise_space_on == "features":
            Uy, Ud = Ud, ops.gaussian_smooth_periodic(Ud, float(denoise_space_sigma))
        else:
            raise ValueError("denoise_space_on must be 'features' or 'all'")
    rng = np.random.default_rng(seed)  # ks2d:1470
    info = {}
    exact = None
    kd = dict(dialect=L.FD_KS_PERIODIC, library=lib)
    if Uy is not None and method not in ("blockwise", "pointwise"):
        raise NotImplementedError("two-stack denoising is served for method='blockwise' and 'pointwise'")
    if method == "blockwise":
        bt, b0, b1 = block
        n_rows = -(-(T - 1) // bt) * -(-A0 // b0) * -(-A1 // b1)
        fold, _ = split_folds(n_rows, rng)
        stats, bad = ops.fd_lib_gram(Ud, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=lib, block=block,
                                     fold_of_row=fold, n_folds=2, variant=variant, return_nonfinite=True, Uy=Uy)
        bad = bad.cpu()
        if int(bad[1]):
            raise ValueError(f"{int(bad[1])} block rows carry a fold id outside [0, 2)")
        if int(bad[0]):
            # non-finite rows renumber the reference's permutation: redo through materialised block rows
            # (pg_fd_block_rows; the rows are 1 / (bt b0 b1) of the field, the drop is host NumPy like the split itself)
            if Uy is not None:
                raise NotImplementedError("non-finite rows together with two-stack denoising")
            rows = _np(ops.fd_block_rows(Ud, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=lib, block=block))
            rows = rows[np.isfinite(rows).all(axis=1)]
            rng = np.random.default_rng(seed)
            fold, _ = split_folds(rows.shape[0], rng)
            Xr, yr, fold_r = ops._dev(np.ascontiguousarray(rows[:, 1:])), ops._dev(np.ascontiguousarray(rows[:, 0])), fold
            stats = ops.rows_gram(Xr, yr, fold_of_row=fold, n_folds=2)[0]
            n_rows = rows.shape[0]
            exact = lambda C: ops.rows_residual_ss(Xr, yr, C, fold_of_row=fold_r, eval_fold=1)   # noqa: E731
        elif Uy is None:
            exact = lambda C: ops.fd_residual_ss(Ud, dx, dy, DT, C, block=block, fold_of_row=fold, n_folds=2,  # noqa: E731
                                                 eval_fold=1, **kd)
        info["X_shape"] = (n_rows, len(names))
    elif method in ("blockwise_left_right", "blockwise_top_bottom"):
        fold = spatial_fold_of_row(T - 1, A0, A1, block, split=method[len("blockwise_"):])
        stats = ops.fd_lib_gram(Ud, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=lib, block=block, fold_of_row=fold,
                                n_folds=2, variant=variant)
        exact = lambda C: ops.fd_residual_ss(Ud, dx, dy, DT, C, block=block, fold_of_row=fold, n_folds=2, eval_fold=1, **kd)  # noqa: E731
        info["X_shape"] = (fold.size, len(names))
    elif method == "pointwise":
        n_total = (T - 1) * A0 * A1
        flat_idx = rng.choice(n_total, size=int(min(n_sample, n_total)), replace=False)
        X, y = ops.fd_gather_rows(Ud, dx, dy, DT, flat_idx, dialect=L.FD_KS_PERIODIC, library=lib, Uy=Uy)
        # ks2d:1633-1636 drops non-finite rows BEFORE the split: rows_gram's statistics turn non-finite when one is
        # present, which is the only case that needs the (host-side, 50 000 x p) filter
        probe = _np(ops.rows_gram(X, y))[0, 0]
        if not np.isfinite(probe).all():
            Xh, yh = _np(X), _np(y)
            ok = np.isfinite(Xh).all(axis=1) & np.isfinite(yh)
            X, y = ops._dev(np.ascontiguousarray(Xh[ok])), ops._dev(np.ascontiguousarray(yh[ok]))
        fold, _ = split_folds(X.shape[0], rng)
        stats = ops.rows_gram(X, y, fold_of_row=fold, n_folds=2)[0]
        exact = lambda C: ops.rows_residual_ss(X, y, C, fold_of_row=fold, eval_fold=1)   # noqa: E731
        info["X_shape"] = tuple(X.shape)
    elif method == "full":
        if fold_of_frame is None:
            raise ValueError("method='full' needs fold_of_frame (time-holdout folds)")
        stats = ops.fd_lib_gram(Ud, dx, dy, DT, dialect=L.FD_KS_PERIODIC, library=lib, block=(1, 1, 1),
                                fold_of_frame=fold_of_frame, n_folds=2, variant=variant)
        exact = lambda C: ops.fd_residual_ss(Ud, dx, dy, DT, C, block=(1, 1, 1), fold_of_frame=fold_of_frame, n_folds=2,  # noqa: E731
                                             eval_fold=1, **kd)
        info["X_shape"] = ((T - 1) * A0 * A1, len(names))
    else:
        raise ValueError(method)
    out = fit_from_stats(stats[0], stats[1], names, alpha=alpha, threshold=threshold, grid_search=grid_search,
                         signs=signs, exact_residuals=exact)
    out.update(info, stats=_np(stats))
    return out
