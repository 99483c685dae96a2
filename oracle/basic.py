"""Oracle: the basic_usage dialect (non-periodic FD, 6-term library, raw-Gram STRidge).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NumPy restatement of
examples/basic_usage.py ("basic").  Layout is ``u[t, y(H), x(W)]``; u_x runs along
the LAST axis (basic:58).
"""

from __future__ import annotations

import numpy as np

TERM_NAMES = ["1", "u", "u_x", "u_y", "lap(u)", "u^2"]  # basic:99


def generate_synthetic_data(n_frames: int = 20, h: int = 50, w: int = 50):
    """basic:13-29: damped travelling sin*cos field (input generator, not hot path)."""
    x = np.linspace(0, 10, w)
    y = np.linspace(0, 10, h)
    t = np.linspace(0, 5, n_frames)
    X, Y = np.meshgrid(x, y)
    tt = t[:, None, None]
    data = np.exp(-0.1 * tt) * np.sin(X[None] - 0.5 * tt) * np.cos(Y[None] - 0.3 * tt)
    return data, x, y, t


def compute_derivatives(u, dx: float, dy: float, dt: float):
    """basic:32-72.  Forward u_t, central interior FD, everything trimmed to
    ``[:-1, 2:-2, 2:-2]``; returns (u_t, u, u_x, u_y, lap_u)."""
    u = np.asarray(u, dtype=np.float64)
    u_t = (u[1:] - u[:-1]) / dt                                             # basic:46-48
    c = u[:-1, 2:-2, 2:-2]
    xm, xp = u[:-1, 2:-2, 1:-3], u[:-1, 2:-2, 3:-1]
    ym, yp = u[:-1, 1:-3, 2:-2], u[:-1, 3:-1, 2:-2]
    u_x = (xp - xm) / (2 * dx)                                              # basic:58
    u_y = (yp - ym) / (2 * dy)                                              # basic:59
    u_xx = (xp - 2 * c + xm) / (dx ** 2)                                    # basic:62
    u_yy = (yp - 2 * c + ym) / (dy ** 2)                                    # basic:63
    return u_t[:, 2:-2, 2:-2], c, u_x, u_y, u_xx + u_yy                     # basic:66-70


def build_library(u, u_x, u_y, lap_u):
    """basic:75-101: Theta (N,6) = [1, u, u_x, u_y, lap, u^2], C-order flatten."""
    f = [np.asarray(a).reshape(-1) for a in (u, u_x, u_y, lap_u)]
    Theta = np.column_stack([np.ones_like(f[0]), f[0], f[1], f[2], f[3], f[0] ** 2])
    return Theta, list(TERM_NAMES)


def stridge_regression(Theta, u_t, alpha: float = 0.01, threshold: float = 0.01, max_iter: int = 10):
    """basic:104-143.  No standardisation.  EVERY iteration restarts from the full ridge
    solve, masks ``|c| < threshold`` and overwrites the active entries with a refit on
    the active columns (so the loop is idempotent); ``max_iter == 0`` returns ones."""
    Theta = np.asarray(Theta)
    u_t = np.asarray(u_t)
    p = Theta.shape[1]
    coef = np.ones(p)
    for _ in range(max_iter):
        coef = np.linalg.solve(Theta.T @ Theta + alpha * np.eye(p), Theta.T @ u_t)
        mask = np.abs(coef) < threshold
        coef[mask] = 0
        act = ~mask
        k = int(act.sum())
        if k == 0:
            break
        Ta = Theta[:, act]
        coef[act] = np.linalg.solve(Ta.T @ Ta + alpha * np.eye(k), Ta.T @ u_t)
    return coef
