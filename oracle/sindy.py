"""Oracle: scripts/patch_based_sindy.py ("sindy"), the per-patch discovery of sindy:226-366 and the ensemble of
sindy:450-470, registration_method="none".

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Restates the reference INCLUDING its feature scramble: build_library
(sindy:247-270) returns ``np.column_stack(terms)`` of eleven (h, w) arrays, i.e. an (h, 11 w) array, which
discover_pde_for_patch views as (h, w, 11) (sindy:327-329).  Pinned by tests/golden/sindy.npz, produced by running the
unmodified class on synthetic images (make_golden.golden_sindy).
"""

from __future__ import annotations

import numpy as np

from . import patch as OP


def compute_derivatives(u, dx, dy):
    """sindy:226-234."""
    ux = (np.roll(u, -1, axis=1) - np.roll(u, 1, axis=1)) / (2 * dx)
    uy = (np.roll(u, -1, axis=0) - np.roll(u, 1, axis=0)) / (2 * dy)
    uxx = (np.roll(u, -1, axis=1) - 2 * u + np.roll(u, 1, axis=1)) / (dx ** 2)
    uyy = (np.roll(u, -1, axis=0) - 2 * u + np.roll(u, 1, axis=0)) / (dy ** 2)
    return ux, uy, uxx, uyy


def build_library(u, ux, uy, uxx, uyy):
    """sindy:247-270 (column_stack of 2-D arrays: (h, 11 w))."""
    lap = uxx + uyy
    return np.column_stack([np.ones_like(u), u, ux, uy, uxx, uyy, lap, u ** 2, u * ux, u * uy, u * lap])


def patch_rows(seq, dx, dy, dt, skip_boundary=5, subsample=4, scramble=True):
    """sindy:300-333: rows of one patch sequence."""
    Xs, ys = [], []
    for i in range(1, len(seq) - 1):
        u = seq[i]
        ut = (seq[i + 1] - seq[i - 1]) / (2 * dt)
        lib = build_library(u, *compute_derivatives(u, dx, dy))
        h, w = u.shape
        mask = np.ones((h, w), dtype=bool)
        mask[:skip_boundary, :] = False
        mask[-skip_boundary:, :] = False
        mask[:, :skip_boundary] = False
        mask[:, -skip_boundary:] = False
        if subsample > 1:
            sub = np.zeros_like(mask)
            sub[::subsample, ::subsample] = True
            mask = mask & sub
        idx = np.where(mask)
        lib3 = lib.reshape(h, w, -1) if scramble else np.stack(np.split(lib, 11, axis=1), axis=2)
        Xs.append(lib3[idx])
        ys.append(ut[idx])
    return np.vstack(Xs), np.concatenate(ys)


def fit_rows(X, y, alpha):
    """sindy:335-356: finite filter, StandardScaler, Ridge(fit_intercept=False), coef / scale_, r2 of X @ coeffs."""
    ok = np.isfinite(X).all(axis=1) & np.isfinite(y)
    X, y = X[ok], y[ok]
    if len(y) < 100:
        return None, 0.0
    mean, scale, _ = OP._scaler_fit(X)
    Xs = (X - mean) / scale
    A = Xs.T @ Xs
    A.flat[:: A.shape[0] + 1] += alpha
    from scipy import linalg

    coef = linalg.solve(A, Xs.T @ y, assume_a="pos") / scale
    pred = X @ coef
    r2 = 1.0 - np.sum((y - pred) ** 2) / np.sum((y - y.mean()) ** 2)
    return coef, max(0.0, r2)


def ensemble(images, patch_size, overlap, dx, dy, dt, alpha=0.01, min_patches=5, scramble=True):
    """sindy:368-470."""
    U = np.asarray(images, dtype=np.float64)
    stride = patch_size - overlap
    h, w = U.shape[1:]
    origins = [(y, x) for y in range(0, h - patch_size + 1, stride) for x in range(0, w - patch_size + 1, stride)]
    C, Q = [], []
    for (y0, x0) in origins:
        seq = [f[y0:y0 + patch_size, x0:x0 + patch_size].copy() for f in U]
        c, q = fit_rows(*patch_rows(seq, dx, dy, dt, scramble=scramble), alpha)
        if c is not None and q > -0.5:
            C.append(c)
            Q.append(q)
    if len(C) < min_patches:
        return None, {}
    C, Q = np.array(C), np.array(Q)
    wts = Q / Q.sum()
    ens = np.average(C, axis=0, weights=wts)
    std = np.sqrt(np.average((C - ens) ** 2, axis=0, weights=wts))
    ens[std > np.median(std) * 2] = 0
    return ens, dict(patch_coeffs=C, patch_qualities=Q, coeffs_std=std)
