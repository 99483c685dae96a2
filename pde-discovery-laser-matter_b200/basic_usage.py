"""Drop-in for the hot-path functions of examples/basic_usage.py ("basic").

Layout ``u[t, y(H), x(W)]``; ``u_x`` runs along the LAST axis (basic:58), so the C ABI is
called with d0 = dy (axis H) and d1 = dx (axis W).
"""

from __future__ import annotations

import numpy as np

from . import _lib as L
from . import ops

TERM_NAMES = ["1", "u", "u_x", "u_y", "lap(u)", "u^2"]  # basic:99


def _np(t):
    from . import _xfer

    return _xfer.to_host(t)       # large results: pinned double-buffered staging


def compute_derivatives(u, dx: float, dy: float, dt: float):
    """basic:32-72: forward u_t and interior central differences, trimmed to
    ``[:-1, 2:-2, 2:-2]``; returns (u_t, u, u_x, u_y, lap_u)."""
    u = np.asarray(u, dtype=np.float64)
    if u.ndim != 3:
        raise ValueError("u must be (T, H, W)")
    T, H, W = u.shape
    if T < 2 or H < 5 or W < 5:
        # same shapes NumPy slicing would give for degenerate inputs
        shp = (max(T - 1, 0), max(H - 4, 0), max(W - 4, 0))
        return tuple(np.zeros(shp) for _ in range(5))
    out = _np(ops.fd_terms(u, dy, dx, dt, dialect=L.FD_BASIC_TRIM, library=L.LIB_BASIC))
    return out[0], out[1], out[2], out[3], out[4]


def build_library(u, u_x, u_y, lap_u):
    """basic:75-101: Theta (N,6) = [1, u, u_x, u_y, lap, u^2] (C-order flatten): pg_basic_library_rows."""
    return _np(ops.basic_library_rows(u, u_x, u_y, lap_u)), list(TERM_NAMES)


def stridge_regression(Theta, u_t, alpha: float = 0.01, threshold: float = 0.01, max_iter: int = 10):
    """basic:104-143 on the GPU: raw Gram (pg_rows_gram) -> K3 in the basic_usage dialect."""
    Theta = np.ascontiguousarray(Theta, dtype=np.float64)
    u_t = np.ascontiguousarray(u_t, dtype=np.float64).reshape(-1)
    if Theta.ndim != 2 or Theta.shape[0] != u_t.shape[0]:
        raise ValueError("Theta must be (N, n_terms) and u_t (N,)")
    stats = ops.rows_gram(Theta, u_t)[:, 0]
    out = ops.stridge_batched(stats, Theta.shape[1], dialect=L.STRIDGE_BASIC, alphas=[alpha], thresholds=[threshold],
                              max_iter=int(max_iter))
    return _np(out["coef"])[0, 0, 0]


def fit_from_field(u, dx, dy, dt, *, alpha: float = 0.01, threshold: float = 0.01, max_iter: int = 10,
                   fold_of_frame=None, n_folds: int = 1, variant=L.VARIANT_AUTO):
    """compute_derivatives + build_library + stridge_regression fused: one pass over ``u``
    accumulates the 6-term Gram (K1), K3 solves; Theta is never materialised."""
    stats = ops.fd_lib_gram(u, dy, dx, dt, dialect=L.FD_BASIC_TRIM, library=L.LIB_BASIC, fold_of_frame=fold_of_frame,
                            n_folds=n_folds, variant=variant)
    out = ops.stridge_batched(stats[0], 6, dialect=L.STRIDGE_BASIC, alphas=[alpha], thresholds=[threshold],
                              max_iter=int(max_iter))
    return dict(coef=_np(out["coef"])[0, 0, 0], names=list(TERM_NAMES), stats=_np(stats))
