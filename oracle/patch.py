"""Oracle: the patch dialect (local cubic-polynomial derivatives, per-patch sklearn STRidge).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NumPy restatement of
scripts/patch_based_pde_discovery.py ("patch").  Layout is ``U[t, y, x]``; the script
stores U as float32 (patch:116) and up-casts each neighbourhood (patch:220).

Third-party arithmetic restated here (not under /root/reference): scikit-learn
(requirements.txt:11 ``>=1.0,<2.0``; 1.9.0 in the build image) ``StandardScaler`` and
``Ridge(alpha)`` on a dense matrix, i.e. population-variance scaling with the
constant-feature rule (sklearn preprocessing/_data.py ``_is_constant_feature`` /
``_handle_zeros_in_scale``) followed by centring X and y and a Cholesky solve of
``(Xs^T Xs + alpha I) w = Xs^T (y - ybar)`` (sklearn linear_model/_ridge.py
``_solve_cholesky``).  tests/test_oracle_golden.py checks this restatement against the
installed scikit-learn.
"""

from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

FULL_NAMES = ["1", "u", "u_x", "u_y", "lap(u)", "u^2", "u*u_x", "u*u_y"]  # patch:375
MODEL4_NAMES = ["1", "u", "u_x", "u_y", "lap(u)", "u^2"]                   # patch:373


# --------------------------------------------------------------------------- derivatives
def poly3d_exponents(deg: int):
    """patch:176-182: (a,b,c) = powers of (t,x,y), a+b+c <= deg, a slowest."""
    return [(a, b, c) for a in range(deg + 1) for b in range(deg + 1 - a) for c in range(deg + 1 - a - b)]


def poly3d_design(t, x, y, exps):
    """patch:185-190."""
    return np.column_stack([(t ** a) * (x ** b) * (y ** c) for a, b, c in exps])


def neighbourhood_design(rt: int, rs: int, deg: int, dt: float, dx: float, dy: float):
    """The (constant) design matrix of patch:206-228; neighbour order t, y, x (x fastest)."""
    tt = np.arange(-rt, rt + 1) * dt
    yy = np.arange(-rs, rs + 1) * dy
    xx = np.arange(-rs, rs + 1) * dx
    Tt, Yy, Xx = np.meshgrid(tt, yy, xx, indexing="ij")
    exps = poly3d_exponents(deg)
    return poly3d_design(Tt.ravel(), Xx.ravel(), Yy.ravel(), exps), exps


def local_poly_derivatives(U, t0, y0, x0, rt, rs, deg, dt, dx, dy):
    """patch:193-246: lstsq cubic fit on the (2rt+1)(2rs+1)^2 neighbourhood ->
    (u, u_t, u_x, u_y, u_xx, u_yy) = coefficients (000),(100),(010),(001),2*(020),2*(002)."""
    A, exps = neighbourhood_design(rt, rs, deg, dt, dx, dy)
    vals = U[t0 - rt:t0 + rt + 1, y0 - rs:y0 + rs + 1, x0 - rs:x0 + rs + 1].astype(np.float64)
    coef, *_ = np.linalg.lstsq(A, vals.ravel(), rcond=None)

    def g(a, b, c):
        return float(coef[exps.index((a, b, c))]) if (a, b, c) in exps else 0.0

    return g(0, 0, 0), g(1, 0, 0), g(0, 1, 0), g(0, 0, 1), 2.0 * g(0, 2, 0), 2.0 * g(0, 0, 2)


def poly_stencil(rt: int, rs: int, deg: int, dt: float, dx: float, dy: float):
    """The lstsq fit of patch:231 has a constant design, so the six outputs are a fixed
    linear stencil ``W (6, n_nb)`` = the matching rows of pinv(A) (x2 for u_xx, u_yy).
    Missing exponents (deg < 2) give zero rows, as get_coef's ValueError branch does."""
    A, exps = neighbourhood_design(rt, rs, deg, dt, dx, dy)
    P = np.linalg.pinv(A)
    W = np.zeros((6, A.shape[0]))
    for r, (e, f) in enumerate([((0, 0, 0), 1.0), ((1, 0, 0), 1.0), ((0, 1, 0), 1.0),
                                ((0, 0, 1), 1.0), ((0, 2, 0), 2.0), ((0, 0, 2), 2.0)]):
        if e in exps:
            W[r] = f * P[exps.index(e)]
    return W


@dataclass(frozen=True)
class Library:
    """patch:156-173."""

    names: list

    def feature_vector(self, u, ux, uy, uxx, uyy):
        lap = uxx + uyy
        if self.names == MODEL4_NAMES:
            return np.array([1.0, u, ux, uy, lap, u ** 2])
        return np.array([1.0, u, ux, uy, lap, u ** 2, u * ux, u * uy])


def build_dataset(U, points, rt, rs, deg, dt, dx, dy, lib: Library):
    """patch:263-280 (per-point lstsq; slow, small cases only)."""
    rows, y = [], []
    for t0, y0, x0 in points:
        u0, ut0, ux0, uy0, uxx0, uyy0 = local_poly_derivatives(U, t0, y0, x0, rt, rs, deg, dt, dx, dy)
        rows.append(lib.feature_vector(u0, ux0, uy0, uxx0, uyy0))
        y.append(ut0)
    return np.vstack(rows), np.array(y)


def build_dataset_stencil(U, points, rt, rs, deg, dt, dx, dy, lib: Library, W=None):
    """Same rows as build_dataset via the fixed stencil (vectorised; what K2 computes)."""
    if W is None:
        W = poly_stencil(rt, rs, deg, dt, dx, dy)
    pts = np.asarray(points, dtype=np.int64).reshape(-1, 3)
    ot, oy, ox = np.meshgrid(np.arange(-rt, rt + 1), np.arange(-rs, rs + 1), np.arange(-rs, rs + 1), indexing="ij")
    V = U[pts[:, 0, None] + ot.ravel()[None], pts[:, 1, None] + oy.ravel()[None],
          pts[:, 2, None] + ox.ravel()[None]].astype(np.float64)
    D = V @ W.T  # (n, 6): u, ut, ux, uy, uxx, uyy
    u, ut, ux, uy = D[:, 0], D[:, 1], D[:, 2], D[:, 3]
    lap = D[:, 4] + D[:, 5]
    cols = [np.ones_like(u), u, ux, uy, lap, u ** 2]
    if lib.names != MODEL4_NAMES:
        cols += [u * ux, u * uy]
    return np.column_stack(cols), ut


# --------------------------------------------------------------------------- patches + sampling
def patch_grid(h: int, w: int, patch: int, overlap: int):
    """patch:283-289: top-left coords with stride max(1, patch-overlap)."""
    s = max(1, patch - overlap)
    return [(y0, x0) for y0 in range(0, h - patch + 1, s) for x0 in range(0, w - patch + 1, s)]


def time_split(t_len: int, rt: int, train_frac: float):
    """patch:361-369."""
    t_valid = np.arange(rt, t_len - rt - 1 + 1)
    split = int(math.floor(train_frac * len(t_valid)))
    return t_valid, t_valid[:split], t_valid[split:]


def sample_patch_points(rng, coords, h, w, patch, rs, t_train, t_test, n_s):
    """patch:395-417: the exact RNG draw order, one sequential stream across patches.
    Returns a list of (train_pts (n_s,3), test_pts (n_te,3)) int arrays [t, y, x]; patches
    whose sampling window is empty are skipped (patch:403-404)."""
    out = []
    n_te = max(30, n_s // 3)
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
        out.append((np.stack([ts, ys, xs], 1), np.stack([ts2, ys2, xs2], 1)))
    return out


# --------------------------------------------------------------------------- sklearn-dialect STRidge
def _scaler_fit(X):
    """sklearn StandardScaler.fit on dense float64: mean, population var, scale with the
    constant-feature rule  var <= n*eps*var + (n*mean*eps)^2  ->  scale 1."""
    n = X.shape[0]
    mean = X.mean(axis=0)
    var = X.var(axis=0)
    eps = np.finfo(np.float64).eps
    const = var <= n * eps * var + (n * mean * eps) ** 2
    scale = np.sqrt(var)
    scale[const | (scale < 10 * eps)] = 1.0
    return mean, scale, const


def _ridge_intercept_coef(Xs, y, alpha):
    """sklearn Ridge(alpha, fit_intercept=True).fit(Xs, y).coef_ on dense data (cholesky)."""
    from scipy import linalg

    Xc = Xs - Xs.mean(axis=0)
    yc = y - y.mean()
    A = Xc.T @ Xc
    A.flat[:: A.shape[0] + 1] += alpha
    return linalg.solve(A, Xc.T @ yc, assume_a="pos")


def stridge(X, y, alpha: float = 0.01, threshold: float = 1e-5, max_iter: int = 25):
    """patch:78-98."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    mean, scale, _ = _scaler_fit(X)
    Xs = (X - mean) / scale
    c = _ridge_intercept_coef(Xs, y, alpha)
    for _ in range(max_iter):
        small = np.abs(c) < threshold
        c[small] = 0
        big = ~small
        if big.sum() == 0:
            break
        cb = _ridge_intercept_coef(Xs[:, big], y, alpha)
        c = np.zeros_like(c)
        c[big] = cb
    return c / (scale + 1e-12)


def stability_aggregate(C, threshold: float, stability_freq: float = 0.6):
    """patch:434-443."""
    nonzero = np.abs(C) > threshold
    freq = nonzero.mean(axis=0)
    median = np.median(C, axis=0)
    q25 = np.percentile(C, 25, axis=0)
    q75 = np.percentile(C, 75, axis=0)
    sign_stability = np.mean(np.sign(C) == np.sign(median + 1e-12), axis=0)
    agg = np.where(freq >= float(stability_freq), median, 0.0)
    return dict(freq=freq, median=median, q25=q25, q75=q75, sign_stability=sign_stability, agg=agg)


def run_patches(U, *, rt=2, rs=3, deg=3, patch=21, overlap=10, samples_per_patch=120, train_frac=0.7,
                alpha=0.01, threshold=1e-5, seed=0, model="full", dx=0.1, dy=0.1, dt=1.0,
                use_stencil=True, max_patches=None):
    """patch:351-443: time split, patch grid, sampling, per-patch datasets + STRidge, aggregation."""
    t_len, h, w = U.shape
    _, t_train, t_test = time_split(t_len, rt, train_frac)
    lib = Library(names=MODEL4_NAMES if model == "model4" else FULL_NAMES)
    coords = patch_grid(h, w, patch, overlap)
    rng = np.random.default_rng(seed)
    samples = sample_patch_points(rng, coords, h, w, patch, rs, t_train, t_test, int(samples_per_patch))
    if max_patches is not None:
        samples = samples[:max_patches]
    W = poly_stencil(rt, rs, deg, dt, dx, dy) if use_stencil else None
    C = []
    for tr_pts, _te_pts in samples:
        if use_stencil:
            X, y = build_dataset_stencil(U, tr_pts, rt, rs, deg, dt, dx, dy, lib, W)
        else:
            X, y = build_dataset(U, [tuple(p) for p in tr_pts.tolist()], rt, rs, deg, dt, dx, dy, lib)
        C.append(stridge(X, y, alpha=alpha, threshold=threshold))
    C = np.stack(C)
    out = stability_aggregate(C, threshold)
    out.update(C=C, samples=samples, rng=rng)
    return out


def regression_metrics(y_true, y_pred) -> dict:
    """patch:47-65 (r2 through sklearn.metrics.r2_score's formula 1 - ss_res / ss_tot)."""
    y_true = np.asarray(y_true).ravel()
    y_pred = np.asarray(y_pred).ravel()
    resid = y_true - y_pred
    rmse = float(np.sqrt(np.mean(resid ** 2)))
    y_std = float(np.std(y_true))
    ss_tot = float(np.sum((y_true - y_true.mean()) ** 2))
    return {"r2": float(1.0 - np.sum(resid ** 2) / ss_tot), "rmse": rmse, "mae": float(np.mean(np.abs(resid))),
            "nrmse": float(rmse / (y_std + 1e-12)),
            "corr": float(np.corrcoef(y_true, y_pred)[0, 1]) if y_true.size > 1 else float("nan"),
            "resid_mean": float(np.mean(resid)), "resid_std": float(np.std(resid)),
            "resid_med_abs": float(np.median(np.abs(resid)))}
