"""Oracle: the analyze_results dialect (scripts/analyze_results.py, "ar"): slice-aligned finite differences with a
CENTRAL time difference, the six nested model libraries and scikit-learn STRidge, plus its validation helpers.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  analyze_results.py is module-level code that reads TIFF files
which are not in the tree, so it cannot be imported; tests/golden/make_golden.py executes its own source lines
(:255-278 derivatives / alignment / split_time, :547-566 stridge, :598-624 models, and the helper functions) on a
synthetic float64 stack and stores inputs and outputs in tests/golden/analyze.npz, which pins this restatement.

The script's stacks are float32 (ar:213: cv2 images; NumPy keeps float32 through ``/ (2*dx)`` and ``u**2``), so the
reference's own derivative / library arithmetic runs in float32 there; this restatement (and the GPU path) works in
float64 on whatever dtype it is given up-cast, which is what the goldens pin.
"""

from __future__ import annotations

import numpy as np

from . import patch as OP

FULL_NAMES = ["1", "u", "u_x", "u_y", "u_xx", "u_yy", "lap(u)", "u^2", "u*u_x", "u*u_y", "u^3", "u_x^2", "u_y^2"]  # ar:622
MODELS = {                                                                                              # ar:598-624
    "Model 1: Diffusion only": ["1", "u", "lap(u)"],
    "Model 2: Diffusion + Linear Growth": ["1", "u", "lap(u)"],
    "Model 3: + First order spatial": ["1", "u", "u_x", "u_y", "lap(u)"],
    "Model 4: + Nonlinear (u^2)": ["1", "u", "u_x", "u_y", "lap(u)", "u^2"],
    "Model 5: + Advection (u*grad(u))": ["1", "u", "u_x", "u_y", "lap(u)", "u^2", "u*u_x", "u*u_y"],
    "Model 6: Full (original)": list(FULL_NAMES),
}


def split_time(t_len: int, train_frac: float):
    """ar:189-194."""
    if not (0.4 <= train_frac <= 0.9):
        raise ValueError("TRAIN_FRAC should be in [0.4, 0.9]")
    split = int(np.floor(train_frac * t_len))
    split = max(1, min(t_len - 1, split))
    return slice(0, split), slice(split, t_len)


def derivatives(U, dx: float, dy: float, dt: float):
    """ar:257-274: slices, then everything cropped to the common origin [:min_t, :min_h, :min_w]."""
    U = np.asarray(U)
    u_x = (U[:, :, 2:] - U[:, :, :-2]) / (2 * dx)
    u_y = (U[:, 2:, :] - U[:, :-2, :]) / (2 * dy)
    u_xx = (U[:, :, 2:] - 2 * U[:, :, 1:-1] + U[:, :, :-2]) / (dx ** 2)
    u_yy = (U[:, 2:, :] - 2 * U[:, 1:-1, :] + U[:, :-2, :]) / (dy ** 2)
    u_t = (U[2:, :, :] - U[:-2, :, :]) / (2 * dt)
    T, H, W = U.shape[0] - 2, U.shape[1] - 2, U.shape[2] - 2
    c = lambda a: a[:T, :H, :W]  # noqa: E731
    u, u_x, u_y, u_xx, u_yy, u_t = c(U), c(u_x), c(u_y), c(u_xx), c(u_yy), c(u_t)
    return dict(u=u, u_x=u_x, u_y=u_y, u_xx=u_xx, u_yy=u_yy, u_t=u_t, lap=u_xx + u_yy)


def library_terms(d):
    """ar:598-624: name -> term array (Model 6's 13 terms; the other models are subsets)."""
    u, u_x, u_y = d["u"], d["u_x"], d["u_y"]
    return {"1": np.ones_like(u), "u": u, "u_x": u_x, "u_y": u_y, "u_xx": d["u_xx"], "u_yy": d["u_yy"], "lap(u)": d["lap"],
            "u^2": u ** 2, "u*u_x": u * u_x, "u*u_y": u * u_y, "u^3": u ** 3, "u_x^2": u_x ** 2, "u_y^2": u_y ** 2}


def rows(d, names, sl):
    """ar:629-632: X = column_stack of term[sl].ravel(), y = u_t[sl].ravel()."""
    terms = library_terms(d)
    return np.column_stack([terms[n][sl].ravel() for n in names]), d["u_t"][sl].ravel()


def stridge(X, y, alpha: float = 0.01, threshold: float = 1e-5, max_iter: int = 20):
    """ar:547-566: the patch dialect's StandardScaler + Ridge STRidge with 20 iterations and ``coeffs / scaler.scale_``
    (no + 1e-12).  Returns (coeffs, scale)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    mean, scale, _ = OP._scaler_fit(X)
    Xs = (X - mean) / scale
    coeffs = OP._ridge_intercept_coef(Xs, y, alpha)
    for _ in range(int(max_iter)):
        small = np.abs(coeffs) < threshold
        coeffs[small] = 0
        big = ~small
        if big.sum() == 0:
            break
        cb = OP._ridge_intercept_coef(Xs[:, big], y, alpha)
        coeffs = np.zeros_like(coeffs)
        coeffs[big] = cb
    return coeffs / scale, scale


def one_step_prediction_rmse(u_field, ut_pred, dt: float = 1.0, spatial_mask=None) -> float:
    """ar:150-187."""
    t_max = min(u_field.shape[0] - 1, ut_pred.shape[0])
    if t_max <= 0:
        return float("nan")
    err = (u_field[1:t_max + 1] - (u_field[:t_max] + dt * ut_pred[:t_max])) ** 2
    if spatial_mask is not None:
        err = err[np.broadcast_to(np.asarray(spatial_mask, dtype=bool), err.shape)]
    return float(np.sqrt(np.mean(err)))


def derivs_2d(f, dx: float, dy: float):
    """ar:300-313: same-grid derivatives through reflect padding (the ROLLOUT's stencils, not the fit's)."""
    f = np.asarray(f, dtype=np.float64)
    fp = np.pad(f, ((1, 1), (1, 1)), mode="reflect")
    u_x = (fp[1:-1, 2:] - fp[1:-1, :-2]) / (2.0 * dx)
    u_y = (fp[2:, 1:-1] - fp[:-2, 1:-1]) / (2.0 * dy)
    u_xx = (fp[1:-1, 2:] - 2.0 * fp[1:-1, 1:-1] + fp[1:-1, :-2]) / (dx ** 2)
    u_yy = (fp[2:, 1:-1] - 2.0 * fp[1:-1, 1:-1] + fp[:-2, 1:-1]) / (dy ** 2)
    return u_x, u_y, u_xx, u_yy, u_xx + u_yy


def ut_from_pde(u2d, terms, coeffs, dx: float, dy: float):
    """ar:316-345."""
    u_x, u_y, u_xx, u_yy, lap = derivs_2d(u2d, dx, dy)
    u = u2d.astype(np.float64)
    tm = {"1": np.ones_like(u), "u": u, "u_x": u_x, "u_y": u_y, "u_xx": u_xx, "u_yy": u_yy, "lap(u)": lap, "u^2": u ** 2,
          "u^3": u ** 3, "u*u_x": u * u_x, "u*u_y": u * u_y, "u_x^2": u_x ** 2, "u_y^2": u_y ** 2}
    out = np.zeros_like(u)
    for name, c in zip(terms, coeffs):
        if abs(float(c)) < 1e-12:
            continue
        out += float(c) * tm[name]
    return out


def rollout_k_rmse(u_true, terms, coeffs, k: int, time_slice, dx: float, dy: float, dt: float, spatial_mask=None):
    """ar:348-395: k explicit-Euler steps from every start frame of the slice; rmse and nrmse over all of them."""
    if k <= 0:
        return {"rmse": float("nan"), "nrmse": float("nan")}
    t0 = time_slice.start or 0
    t1 = min(time_slice.stop or u_true.shape[0], u_true.shape[0])
    if t1 - t0 <= k:
        return {"rmse": float("nan"), "nrmse": float("nan")}
    errs, trues = [], []
    for t in range(t0, t1 - k):
        up = u_true[t].astype(np.float64)
        for _ in range(k):
            up = up + dt * ut_from_pde(up, terms, coeffs, dx, dy)
        target = u_true[t + k].astype(np.float64)
        diff = target - up
        if spatial_mask is not None:
            m = np.asarray(spatial_mask, dtype=bool)
            diff, target = diff[m], target[m]
        errs.append(diff.ravel())
        trues.append(target.ravel())
    e, yv = np.concatenate(errs), np.concatenate(trues)
    rmse = float(np.sqrt(np.mean(e ** 2)))
    return {"rmse": rmse, "nrmse": float(rmse / (float(np.std(yv)) + 1e-12))}


def fit_models(U, dx: float, dy: float, dt: float, train_frac: float = 0.7, alpha: float = 0.01, threshold: float = 1e-5):
    """ar:626-640 for every model: time split, rows, STRidge, train / test metrics."""
    d = derivatives(U, dx, dy, dt)
    tr, te = split_time(d["u"].shape[0], train_frac)
    out = {}
    for name, names in MODELS.items():
        Xtr, ytr = rows(d, names, tr)
        Xte, yte = rows(d, names, te)
        c, scale = stridge(Xtr, ytr, alpha=alpha, threshold=threshold)
        out[name] = dict(names=names, coeffs=c, scale=scale, train=OP.regression_metrics(ytr, Xtr @ c),
                         test=OP.regression_metrics(yte, Xte @ c))
    return out
