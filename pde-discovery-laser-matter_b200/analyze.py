"""Drop-in for the hot path of scripts/analyze_results.py ("ar"), the reference's real-data pipeline.

The script is module-level code, so its boundary is a handful of helper functions (``stridge``, ``split_time``,
``regression_metrics``, ``one_step_prediction_rmse``, ``rollout_k_rmse``) plus the inline block that differentiates
the stack and fits six nested models (ar:255-278, 598-640).  Here:

* ``fit_models`` is that inline block fused: ONE pass of K1 in the slice-central dialect (PG_FD_SLICE_CENTRAL: every
  difference is a pair / triple of slices cropped to a common origin, central time difference, ar:257-274) accumulates
  the statistics of the 13-term library of "Model 6" for the train and the test time segment (``split_time``,
  ar:189-194); Models 1-5 are column subsets, so their statistics are entries of the same vector, and one K3 launch
  (scikit-learn dialect, ``coeffs / scaler.scale_`` without the 1e-12 of the patch script, 20 iterations) fits all six.
* the helper functions keep their signatures and run on the GPU.

Layout ``U[t, y, x]`` (x = last axis); the script's stacks are float32 and NumPy keeps them float32 through the
differences, this path computes in float64 on the up-cast values (see oracle/analyze.py).
"""

from __future__ import annotations

import numpy as np

from . import _lib as L
from . import ops
from .patch import regression_metrics  # ar:137-156 is the same function as patch:47-65  # noqa: F401

FULL_NAMES = ["1", "u", "u_x", "u_y", "u_xx", "u_yy", "lap(u)", "u^2", "u*u_x", "u*u_y", "u^3", "u_x^2", "u_y^2"]  # ar:622
MODELS = {                                                                                              # ar:598-624
    "Model 1: Diffusion only": ["1", "u", "lap(u)"],
    "Model 2: Diffusion + Linear Growth": ["1", "u", "lap(u)"],
    "Model 3: + First order spatial": ["1", "u", "u_x", "u_y", "lap(u)"],
    "Model 4: + Nonlinear (u^2)": ["1", "u", "u_x", "u_y", "lap(u)", "u^2"],
    "Model 5: + Advection (u*grad(u))": ["1", "u", "u_x", "u_y", "lap(u)", "u^2", "u*u_x", "u*u_y"],
    "Model 6: Full (original)": list(FULL_NAMES),
}


def _np(t):
    from . import _xfer

    return _xfer.to_host(t)


def split_time(t_len: int, train_frac: float):
    """ar:189-194."""
    if not (0.4 <= train_frac <= 0.9):
        raise ValueError("TRAIN_FRAC should be in [0.4, 0.9]")
    split = int(np.floor(train_frac * t_len))
    split = max(1, min(t_len - 1, split))
    return slice(0, split), slice(split, t_len)


class Scaler:
    """What the script uses of the StandardScaler that its stridge returns (ar:566, 578): ``scale_`` and ``mean_``."""

    def __init__(self, mean, scale):
        self.mean_, self.scale_ = mean, scale


def stridge(X, y, alpha=0.01, threshold=1e-5, max_iter=20):
    """ar:547-566: rows -> statistics (shifted by the first row) -> K3, scikit-learn dialect without the 1e-12 in the
    unscaling.  Returns (coeffs, scaler) like the script."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.ascontiguousarray(y, dtype=np.float64)
    if X.ndim != 2 or y.shape != (X.shape[0],):
        raise ValueError("X must be (n, p) and y (n,)")
    shift = X[:1].copy()
    stats, mm = ops.rows_gram(X, y, shift=shift, want_minmax=True)
    out = ops.stridge_batched(stats[:, 0], X.shape[1], dialect=L.STRIDGE_SKLEARN, flags=L.STRIDGE_NO_EPS, alphas=[alpha],
                              thresholds=[threshold], max_iter=int(max_iter), colminmax=mm[:, 0], shift=shift)
    s, m = _np(stats)[0, 0], _np(mm)[0, 0]
    return _np(out["coef"])[0, 0, 0], Scaler(*scaler_from_stats(s, X.shape[1], shift[0], m))


def scaler_from_stats(s, p, shift=None, colminmax=None):
    """StandardScaler's mean_ / scale_ (population std, constant columns -> 1) from a statistics vector."""
    n = s[0]
    mu = s[3:3 + p] / n
    diag = np.array([s[3 + 2 * p + i * p - (i * (i - 1)) // 2] for i in range(p)])
    var = np.maximum(diag / n - mu * mu, 0.0)
    mean = mu + (0.0 if shift is None else shift)
    eps = np.finfo(np.float64).eps
    const = var <= n * eps * var + (n * mean * eps) ** 2
    if colminmax is not None:
        const |= colminmax[0] == colminmax[1]
    scale = np.sqrt(var)
    scale[const | (scale < 10 * eps)] = 1.0
    return mean, scale


def sub_stats(stats, cols, p_full):
    """Statistics vector of a column subset, gathered from the full one (a few dozen entries: host index arithmetic)."""
    s = np.asarray(stats, dtype=np.float64)
    cols = list(cols)
    q = len(cols)
    gidx = lambda i, j: 3 + 2 * p_full + i * p_full - (i * (i - 1)) // 2 + (j - i)  # noqa: E731
    out = np.empty(L.stats_len(q))
    out[:3] = s[:3]
    out[3:3 + q] = s[3 + np.array(cols)]
    out[3 + q:3 + 2 * q] = s[3 + p_full + np.array(cols)]
    k = 3 + 2 * q
    for a in range(q):
        for b in range(a, q):
            i, j = sorted((cols[a], cols[b]))
            out[k] = s[gidx(i, j)]
            k += 1
    return out


def model_stats(U, dx, dy, dt, train_frac=0.7, spatial_mask=None):
    """One K1 pass (generic kernel, reference arithmetic): statistics [2][S(13)] of the train / test time segments of
    ar:278 (``spatial_mask``: optional (H-2, W-2) boolean TRAIN region, ar:282-299; rows outside it go to the test fold)."""
    Ud = ops.field(U)
    Tr, R0, R1 = ops.row_space(Ud.shape, L.FD_SLICE_CENTRAL)
    if Tr < 2:
        raise ValueError("need at least 4 frames (two row frames after the central time difference)")
    tr, te = split_time(Tr, train_frac)
    fof = (np.arange(Tr) >= tr.stop).astype(np.int32)
    kw = dict(dialect=L.FD_SLICE_CENTRAL, library=L.LIB_AR_FULL, n_folds=2)
    if spatial_mask is None:
        return ops.fd_lib_gram(Ud, dy, dx, dt, fold_of_frame=fof, **kw), tr, te
    m = np.asarray(spatial_mask, dtype=bool)
    if m.shape != (R0, R1):
        raise ValueError(f"spatial_mask shape {m.shape} does not match the aligned field {(R0, R1)}")
    fold = np.broadcast_to(np.where(m, 0, 1).astype(np.uint8), (Tr, R0, R1)).reshape(-1)
    return ops.fd_lib_gram(Ud, dy, dx, dt, fold_of_row=fold, **kw), tr, te


def fit_models(U, dx=0.1, dy=0.1, dt=1.0, *, train_frac=0.7, alpha=0.01, threshold=1e-5, max_iter=20, models=None,
               spatial_mask=None):
    """ar:255-278 + 598-640 fused: derivatives, alignment, time split, the six libraries and their STRidge fits.
    Returns {model name: dict(names, coeffs, scale, n_active, train=dict(r2, rmse), test=dict(r2, rmse))}."""
    models = MODELS if models is None else models
    stats, tr, te = model_stats(U, dx, dy, dt, train_frac, spatial_mask)
    S = _np(stats)
    p_full = len(FULL_NAMES)
    out = {}
    for name, names in models.items():
        cols = [FULL_NAMES.index(n) for n in names]
        s_tr, s_te = sub_stats(S[0], cols, p_full), sub_stats(S[1], cols, p_full)
        q = len(cols)
        const_cols = [k for k, n in enumerate(names) if n == "1"]
        fit = ops.stridge_batched(np.stack([s_tr, s_tr]), q, dialect=L.STRIDGE_SKLEARN, flags=L.STRIDGE_NO_EPS, alphas=[alpha],
                                  thresholds=[threshold], max_iter=int(max_iter), const_cols=const_cols,
                                  eval_stats=np.stack([s_tr, s_te]))
        c = _np(fit["coef"])[0, 0, 0]
        met = _np(fit["metrics"])[:, 0, 0]
        _, scale = scaler_from_stats(s_tr, q)
        scale[const_cols] = 1.0
        out[name] = dict(names=list(names), coeffs=c, scale=scale, n_active=int(np.sum(np.abs(c) > 1e-5)),   # ar:660
                         train=dict(r2=float(met[0, 0]), rmse=float(met[0, 1])),
                         test=dict(r2=float(met[1, 0]), rmse=float(met[1, 1])))
    return out


def one_step_prediction_rmse(u_field, ut_pred, dt=1.0, spatial_mask=None):
    """ar:150-187: u(t+1) ~ u(t) + dt * u_t_pred(t)."""
    if spatial_mask is not None:
        m = np.asarray(spatial_mask, dtype=bool)
        if m.ndim != 2:
            raise ValueError("spatial_mask must be 2D (HxW)")
        if m.shape != tuple(np.asarray(u_field).shape[1:]):
            raise ValueError(f"spatial_mask shape {m.shape} does not match field shape {tuple(np.asarray(u_field).shape[1:])}")
    s = ops.one_step_sums(u_field, ut_pred, dt, spatial_mask)
    if s is None or s[1] == 0:
        return float("nan")
    return float(np.sqrt(s[0] / s[1]))


def rollout_k_rmse(u_true, model_terms, model_coeffs, k, time_slice, spatial_mask=None, *, dx=0.1, dy=0.1, dt=1.0):
    """ar:348-395: k-step explicit-Euler rollout RMSE over every start time of the slice (the script takes dx, dy, dt
    from its module globals: keyword arguments here)."""
    if k <= 0:
        return {"rmse": float("nan"), "nrmse": float("nan")}
    u_true = np.asarray(u_true)
    t0 = time_slice.start or 0
    t1 = min(time_slice.stop or u_true.shape[0], u_true.shape[0])
    if t1 - t0 <= k:
        return {"rmse": float("nan"), "nrmse": float("nan")}
    for n in model_terms:
        if n not in FULL_NAMES:
            raise KeyError(f"Rollout: unsupported term '{n}'")
    ids = [FULL_NAMES.index(n) for n in model_terms]
    s = ops.ar_rollout_sums(np.ascontiguousarray(u_true, dtype=np.float64), dy, dx, dt, ids, model_coeffs, int(k), t0, t1,
                            spatial_mask)
    n = s[3]
    rmse = float(np.sqrt(s[0] / n))
    std = float(np.sqrt(max(s[2] / n - (s[1] / n) ** 2, 0.0)))
    return {"rmse": rmse, "nrmse": float(rmse / (std + 1e-12))}
