"""Drop-in for the hot-path functions of scripts/patch_based_pde_discovery.py ("patch").

Layout ``U[t, y, x]`` (float32 as the script stores it, patch:116, or float64).  The lstsq
fit of patch:231 has a constant design matrix, so the derivative estimate is a fixed
(2rt+1)(2rs+1)^2-tap stencil ``W6 = pinv(A)[rows]`` applied by kernel K2; per-patch STRidge
(StandardScaler + Ridge, patch:78-98) is K3 in the scikit-learn dialect.
"""

from __future__ import annotations

import math
from dataclasses import dataclass
from functools import lru_cache

import numpy as np

from . import _lib as L
from . import ops

FULL_NAMES = ["1", "u", "u_x", "u_y", "lap(u)", "u^2", "u*u_x", "u*u_y"]  # patch:375
MODEL4_NAMES = ["1", "u", "u_x", "u_y", "lap(u)", "u^2"]                   # patch:373


def _np(t):
    from . import _xfer

    return _xfer.to_host(t)       # large results: pinned double-buffered staging


def _poly3d_exponents(deg: int):
    """patch:176-182."""
    return [(a, b, c) for a in range(deg + 1) for b in range(deg + 1 - a) for c in range(deg + 1 - a - b)]


@lru_cache(maxsize=32)
def poly_stencil(rt: int, rs: int, deg: int, dt: float, dx: float, dy: float):
    """The constant stencil equivalent to patch:206-246: rows (u, u_t, u_x, u_y, u_xx, u_yy) of
    pinv(design) with the factor 2 on the second derivatives; neighbour order t, y, x.
    A p x 245 pseudo-inverse computed once on the host (setup, like 1/dx^2), not per point."""
    tt = np.arange(-rt, rt + 1) * dt
    yy = np.arange(-rs, rs + 1) * dy
    xx = np.arange(-rs, rs + 1) * dx
    Tt, Yy, Xx = np.meshgrid(tt, yy, xx, indexing="ij")
    exps = _poly3d_exponents(deg)
    A = np.column_stack([(Tt.ravel() ** a) * (Xx.ravel() ** b) * (Yy.ravel() ** c) for a, b, c in exps])
    P = np.linalg.pinv(A)
    W = np.zeros((6, A.shape[0]))
    for r, (e, f) in enumerate([((0, 0, 0), 1.0), ((1, 0, 0), 1.0), ((0, 1, 0), 1.0), ((0, 0, 1), 1.0),
                                ((0, 2, 0), 2.0), ((0, 0, 2), 2.0)]):
        if e in exps:  # get_coef returns 0.0 for a missing exponent (patch:233-238)
            W[r] = f * P[exps.index(e)]
    return W


@dataclass(frozen=True)
class Library:
    """patch:156-173."""

    names: list

    @property
    def library_id(self) -> int:
        return L.LIB_PATCH_MODEL4 if self.names == MODEL4_NAMES else L.LIB_PATCH_FULL

    def feature_vector(self, u, ux, uy, uxx, uyy):
        lap = uxx + uyy
        if self.names == MODEL4_NAMES:
            return np.array([1.0, u, ux, uy, lap, u ** 2], dtype=np.float64)
        return np.array([1.0, u, ux, uy, lap, u ** 2, u * ux, u * uy], dtype=np.float64)


def _check_points(U, pts, rt, rs):
    T, H, W = U.shape
    pts = np.asarray(pts, dtype=np.int64).reshape(-1, 3)
    if len(pts) and ((pts[:, 0] - rt).min() < 0 or (pts[:, 0] + rt).max() >= T or (pts[:, 1] - rs).min() < 0
                     or (pts[:, 1] + rs).max() >= H or (pts[:, 2] - rs).min() < 0 or (pts[:, 2] + rs).max() >= W):
        raise IndexError("derivative neighbourhood leaves the stack")  # NumPy raises IndexError at patch:220
    return pts.astype(np.int32)


def local_poly_derivatives(U, t0, y0, x0, rt, rs, deg, dt, dx, dy):
    """patch:193-246: (u, u_t, u_x, u_y, u_xx, u_yy) at one point."""
    pts = _check_points(U, [(t0, y0, x0)], rt, rs)
    D, _ = ops.poly_rows(U, pts, poly_stencil(rt, rs, deg, float(dt), float(dx), float(dy)), rt, rs,
                         library=L.LIB_PATCH_DERIVS)
    return tuple(float(v) for v in _np(D)[0])


def build_dataset(U, points, rt, rs, deg, dt, dx, dy, lib: Library):
    """patch:263-280: library rows and u_t for a list of (t, y, x) points, one K2 launch."""
    pts = _check_points(U, points, rt, rs)
    X, y = ops.poly_rows(U, pts, poly_stencil(rt, rs, deg, float(dt), float(dx), float(dy)), rt, rs,
                         library=lib.library_id)
    return _np(X), _np(y)


def patch_grid(h: int, w: int, patch: int, overlap: int):
    """patch:283-289."""
    stride = max(1, patch - overlap)
    return [(y0, x0) for y0 in range(0, h - patch + 1, stride) for x0 in range(0, w - patch + 1, stride)]


def stridge(X, y, alpha: float = 0.01, threshold: float = 1e-5, max_iter: int = 25):
    """patch:78-98 on the GPU: rows -> statistics (shifted by the first row so the centred Gram
    does not cancel) -> K3 in the scikit-learn dialect."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if X.ndim != 2 or y.shape != (X.shape[0],):
        raise ValueError("X must be (n, p) and y (n,)")
    shift = X[:1].copy()
    stats, mm = ops.rows_gram(X, y, shift=shift, want_minmax=True)
    out = ops.stridge_batched(stats[:, 0], X.shape[1], dialect=L.STRIDGE_SKLEARN, alphas=[alpha], thresholds=[threshold],
                              max_iter=int(max_iter), colminmax=mm[:, 0], shift=shift)
    return _np(out["coef"])[0, 0, 0]


def regression_metrics(y_true, y_pred) -> dict:
    """patch:47-65 from one two-pass reduction on the GPU (the median of |resid| via torch, on the device)."""
    torch = L.torch_cuda()
    yt = ops._dev(np.asarray(y_true, dtype=np.float64).ravel(), torch.float64)
    yp = ops._dev(np.asarray(y_pred, dtype=np.float64).ravel(), torch.float64)
    s, n = ops.fit_metric_sums(yt, yp)
    rmse = float(np.sqrt(s[1] / n))
    y_std = float(np.sqrt(s[5] / n))
    corr = float(s[7] / np.sqrt(s[5] * s[6])) if n > 1 else float("nan")
    # sklearn.metrics.r2_score (patch:38,57): no guard; a constant y_true scores 1 for a perfect fit, else 0
    r2 = float(1.0 - s[1] / s[5]) if s[5] > 0 else (1.0 if s[1] == 0 else 0.0)
    return {"r2": r2, "rmse": rmse, "mae": float(s[2] / n), "nrmse": float(rmse / (y_std + 1e-12)),
            "corr": corr, "resid_mean": float(s[0] / n), "resid_std": float(np.sqrt(s[8] / n)),
            "resid_med_abs": float(np.median((yt - yp).abs().cpu().numpy()))}


def _metrics_from_sums(s, n, resid_abs_median):
    """patch:47-65 from the sums of pg_fit_metrics / pg_rows_metrics_batched (arrays over a leading batch axis)."""
    s = np.asarray(s, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        rmse = np.sqrt(s[..., 1] / n)
        y_std = np.sqrt(s[..., 5] / n)
        r2 = np.where(s[..., 5] > 0, 1.0 - s[..., 1] / s[..., 5], np.where(s[..., 1] == 0, 1.0, 0.0))
        corr = s[..., 7] / np.sqrt(s[..., 5] * s[..., 6]) if n > 1 else np.full(s.shape[:-1], np.nan)
    return {"r2": r2, "rmse": rmse, "mae": s[..., 2] / n, "nrmse": rmse / (y_std + 1e-12), "corr": corr,
            "resid_mean": s[..., 0] / n, "resid_std": np.sqrt(s[..., 8] / n), "resid_med_abs": resid_abs_median}


def gaussian_filter(stack, sigma):
    """``np.array([gaussian_filter(img, sigma=sigma) for img in stack])`` of patch:335,343 (scipy.ndimage, reflect
    boundary): bit-identical for float32 and float64 stacks (pg_reflect_conv).  A single 2-D frame works too."""
    a = np.asarray(stack)
    if a.ndim == 2:
        return _np(ops.gaussian_filter_frames(a[None], sigma))[0]
    return _np(ops.gaussian_filter_frames(a, sigma))


def safe_sample_points(rng, t_indices, h, w, rs, n):
    """patch:248-260 (host RNG, draw order kept)."""
    ys = rng.integers(rs, h - rs, size=n)
    xs = rng.integers(rs, w - rs, size=n)
    ts = rng.choice(t_indices, size=n, replace=True)
    return list(zip(ts.tolist(), ys.tolist(), xs.tolist()))


def global_checks(U, agg, rng, *, rt=2, rs=3, deg=3, dt=1.0, dx=0.1, dy=0.1, model="full", train_frac=0.7,
                  n_global=800, n_step=1200):
    """patch:446-465 with the aggregated model: (a) regression_metrics on 800 points sampled over the held-out
    frames, (b) the one-step check on 1200 sampled points, sqrt(mean((u(t+1) - u(t) - dt * u_t_pred)^2)).  ``rng`` is
    the generator the per-patch loop has already advanced (the draws continue its stream, as in main())."""
    t_len, h, w = U.shape
    lib = Library(names=MODEL4_NAMES if model == "model4" else FULL_NAMES)
    t_valid, _, t_test = time_split(t_len, rt, train_frac)
    agg = np.asarray(agg, dtype=np.float64)
    pts_g = safe_sample_points(rng, t_test, h, w, rs, n_global)
    Xg, yg = build_dataset(U, pts_g, rt, rs, deg, dt, dx, dy, lib)
    m_test = regression_metrics(yg, Xg @ agg)
    pts_s = safe_sample_points(rng, t_valid[:-1], h, w, rs, n_step)
    Xs, _ = build_dataset(U, pts_s, rt, rs, deg, dt, dx, dy, lib)
    ut_pred = Xs @ agg
    P = np.asarray(pts_s, dtype=np.int64)
    ok = P[:, 0] + 1 < t_len
    Uh = np.asarray(U)
    du = (Uh[P[ok, 0] + 1, P[ok, 1], P[ok, 2]] - Uh[P[ok, 0], P[ok, 1], P[ok, 2]]).astype(np.float64)   # float32 difference, like float(U[..] - U[..])
    one_step = float(np.sqrt(np.mean((du - dt * ut_pred[ok]) ** 2))) if ok.any() else float("nan")
    return dict(test_metrics=m_test, one_step_rmse=one_step, global_points=pts_g, step_points=pts_s)


# ------------------------------------------------------------------ fused per-patch ensemble (patch:351-443)
def time_split(t_len: int, rt: int, train_frac: float):
    """patch:361-369."""
    t_valid = np.arange(rt, t_len - rt)
    split = int(math.floor(train_frac * len(t_valid)))
    return t_valid, t_valid[:split], t_valid[split:]


def sample_patch_points(rng, coords, h, w, patch, rs, t_train, t_test, n_s):
    """patch:395-417: the reference's RNG draw order (one sequential stream across patches) is
    part of the contract, so this loop stays on the host verbatim; returns int32 arrays
    train [B][n_s][3], test [B][n_te][3] of (t, y, x)."""
    n_te = max(30, n_s // 3)
    tr, te = [], []
    for (y0, x0) in coords:
        ylo, yhi = max(rs, y0 + rs), min(h - rs, y0 + patch - rs)
        xlo, xhi = max(rs, x0 + rs), min(w - rs, x0 + patch - rs)
        if yhi <= ylo or xhi <= xlo:
            continue
        ys = rng.integers(ylo, yhi, size=n_s)
        xs = rng.integers(xlo, xhi, size=n_s)
        ts = rng.choice(t_train, size=n_s, replace=True)
        ys2 = rng.integers(ylo, yhi, size=n_te)
        xs2 = rng.integers(xlo, xhi, size=n_te)
        ts2 = rng.choice(t_test, size=n_te, replace=True)
        tr.append(np.stack([ts, ys, xs], 1))
        te.append(np.stack([ts2, ys2, xs2], 1))
    if not tr:
        return np.zeros((0, n_s, 3), np.int32), np.zeros((0, n_te, 3), np.int32)
    return np.stack(tr).astype(np.int32), np.stack(te).astype(np.int32)


def stability_aggregate(C, threshold: float, stability_freq: float = 0.6):
    """patch:434-443 (a few p-vectors from the (B,p) coefficient table; host)."""
    C = np.asarray(C)
    freq = (np.abs(C) > threshold).mean(axis=0)
    median = np.median(C, axis=0)
    return dict(freq=freq, median=median, q25=np.percentile(C, 25, axis=0), q75=np.percentile(C, 75, axis=0),
                sign_stability=np.mean(np.sign(C) == np.sign(median + 1e-12), axis=0),
                agg=np.where(freq >= float(stability_freq), median, 0.0))


def fit_patches(U, *, rt=2, rs=3, deg=3, patch=21, overlap=10, samples_per_patch=120, train_frac=0.7, alpha=0.01,
                threshold=1e-5, seed=0, model="full", dx=0.1, dy=0.1, dt=1.0, max_iter=25, train_pts=None, test_pts=None,
                with_metrics=True):
    """The per-patch loop of main() (patch:351-443) as batched launches: K2 over every sampled point of every patch,
    rows -> per-patch statistics, K3 over all patches, and (``with_metrics``) the per-patch train / test
    regression_metrics of patch:425-429 from one more K2 launch over the test points.  ``rng`` in the result is the
    generator after the loop's draws (global_checks continues its stream)."""
    torch = L.torch_cuda()
    t_len, h, w = U.shape
    lib = Library(names=MODEL4_NAMES if model == "model4" else FULL_NAMES)
    p = len(lib.names)
    rng = None
    if train_pts is None:
        _, t_train, t_test = time_split(t_len, rt, train_frac)
        coords = patch_grid(h, w, patch, overlap)
        rng = np.random.default_rng(seed)
        train_pts, test_pts = sample_patch_points(rng, coords, h, w, patch, rs, t_train, t_test, int(samples_per_patch))
    B, n_s = train_pts.shape[:2]
    W6 = poly_stencil(rt, rs, deg, float(dt), float(dx), float(dy))
    X, y = ops.poly_rows(U, _check_points(U, train_pts, rt, rs), W6, rt, rs, library=lib.library_id)
    X = X.reshape(B, n_s, p)
    y = y.reshape(B, n_s)
    shift = X[:, 0, :].contiguous()
    stats, mm = ops.rows_gram(X, y, shift=shift, want_minmax=True)
    out = ops.stridge_batched(stats[:, 0], p, dialect=L.STRIDGE_SKLEARN, alphas=[alpha], thresholds=[threshold],
                              max_iter=int(max_iter), colminmax=mm[:, 0], shift=shift)
    C = _np(out["coef"])[:, 0, 0, :]
    res = stability_aggregate(C, threshold)
    res.update(C=C, train_pts=train_pts, test_pts=test_pts, names=list(lib.names), rng=rng)
    if with_metrics and B:
        coef_d = out["coef"][:, 0, 0, :].contiguous()
        s_tr, r_tr = ops.rows_metrics_batched(X, y, coef_d, want_resid=True)
        res["train_metrics"] = _metrics_from_sums(_np(s_tr), n_s, np.median(np.abs(_np(r_tr)), axis=1))
        if test_pts is not None and len(test_pts):
            n_te = test_pts.shape[1]
            Xt, yt = ops.poly_rows(U, _check_points(U, test_pts, rt, rs), W6, rt, rs, library=lib.library_id)
            s_te, r_te = ops.rows_metrics_batched(Xt.reshape(B, n_te, p), yt.reshape(B, n_te), coef_d, want_resid=True)
            res["test_metrics"] = _metrics_from_sums(_np(s_te), n_te, np.median(np.abs(_np(r_te)), axis=1))
    return res
