"""The hot path of scripts/patch_based_sindy.py ("sindy"): 256^2 patches with overlap 64, per-patch
periodic-roll differences, an 11-term library, StandardScaler + Ridge(fit_intercept=False) per patch and a
quality-weighted ensemble.

``PatchBasedSINDy`` keeps the reference class's constructor and the two methods on its hot path,
``discover_pde_for_patch`` and ``discover_pde_patch_ensemble`` (registration_method="none": the script's default;
the ECC / optical-flow registrations are OpenCV pre-processing and out of scope).  All patches are processed by
three launches: pg_sindy_rows (rows of every patch), pg_rows_gram (per-patch statistics), pg_stridge_batched (the
ridge solve: centred and scaled X, raw y, no thresholding -- the ks2d dialect with max_iter = 0 and ``coef / scale_``),
plus pg_rows_metrics_batched for the R^2 "quality" of sindy:353-356.

Faithful to a reference quirk: build_library column_stacks 2-D term arrays and discover_pde_for_patch then views the
(h, 11 w) result as (h, w, 11) (sindy:269, 327-329), so the "features" of a pixel are 11 consecutive samples of one or
two terms along its row.  ``scramble=True`` (default) reproduces that; ``scramble=False`` regresses on the eleven terms
at the pixel.
"""

from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from . import ops

TERM_NAMES = ["1", "u", "u_x", "u_y", "u_xx", "u_yy", "∇²u", "u²", "u·u_x", "u·u_y", "u·∇²u"]   # sindy:264-267


def _np(t):
    from . import _xfer

    return _xfer.to_host(t)


def patch_rows(U, origins, patch_size, dx, dy, dt, skip_boundary=5, subsample=4, scramble=True):
    """pg_sindy_rows: (X [B][n][11], y [B][n]) device tensors for the patches at ``origins`` [(y, x), ...]."""
    torch = L.torch_cuda()
    lib = L.load()
    U = ops.field(U)
    T, H, W = U.shape
    org = ops._dev(np.asarray(origins, dtype=np.int32).reshape(-1, 2))
    B = org.shape[0]
    n = C.c_int64()
    args = (int(patch_size), int(skip_boundary), int(subsample), float(dy), float(dx), float(dt), 1 if scramble else 0)
    L.check(lib.pg_sindy_rows(L.ptr(U), T, H, W, L.ptr(org), 0, *args, L.ptr(U), L.ptr(U), C.byref(n), L.stream_ptr()))
    X = torch.empty((B, n.value, 11), dtype=torch.float64, device=U.device)
    y = torch.empty((B, n.value), dtype=torch.float64, device=U.device)
    if B and n.value:
        L.check(lib.pg_sindy_rows(L.ptr(U), T, H, W, L.ptr(org), B, *args, L.ptr(X), L.ptr(y), C.byref(n), L.stream_ptr()))
    return X, y


def fit_patch_rows(X, y, alpha):
    """StandardScaler + Ridge(alpha, fit_intercept=False) + ``coef_ / scale_`` (sindy:343-351) and the R^2 of
    ``X @ coeffs`` against y (sindy:353-355) for B problems: (coeffs [B][11], r2 [B]) NumPy arrays."""
    B, n, p = X.shape
    shift = X[:, 0, :].contiguous()
    stats, mm = ops.rows_gram(X, y, shift=shift, want_minmax=True)
    out = ops.stridge_batched(stats[:, 0], p, dialect=L.STRIDGE_KS, flags=L.STRIDGE_NO_EPS, alphas=[alpha], thresholds=[0.0],
                              max_iter=0, colminmax=mm[:, 0], shift=shift)
    coef = out["coef"][:, 0, 0, :].contiguous()
    s = _np(ops.rows_metrics_batched(X, y, coef))
    with np.errstate(divide="ignore", invalid="ignore"):
        r2 = np.where(s[:, 5] > 0, 1.0 - s[:, 1] / s[:, 5], np.where(s[:, 1] == 0, 1.0, 0.0))     # sklearn.metrics.r2_score
    return _np(coef), r2


class PatchBasedSINDy:
    """sindy:34-60 (constructor) and the discovery methods of sindy:272-366, 368-490."""

    def __init__(self, dt=1.0, dx=1.0, dy=1.0, patch_size=256, overlap=64, scramble=True):
        self.dt, self.dx, self.dy = dt, dx, dy
        self.patch_size, self.overlap, self.stride = patch_size, overlap, patch_size - overlap
        self.images = []
        self.scramble = scramble

    def patch_origins(self, shape):
        """Top-left corners in extract_patches order (sindy:121-137)."""
        h, w = shape
        return [(y, x) for y in range(0, h - self.patch_size + 1, self.stride)
                for x in range(0, w - self.patch_size + 1, self.stride)]

    def _fit(self, U, origins, alpha, skip_boundary, subsample):
        X, y = patch_rows(U, origins, self.patch_size, self.dx, self.dy, self.dt, skip_boundary, subsample, self.scramble)
        B, n, _ = X.shape
        if B == 0 or n == 0:
            return [None] * B, np.zeros(B)
        probe = _np(ops.rows_gram(X, y))[:, 0, :]                  # non-finite rows show up in the statistics
        coeffs, r2 = fit_patch_rows(X, y, alpha)
        res, qual = [], np.zeros(B)
        for b in range(B):
            if not np.isfinite(probe[b]).all():
                # sindy:335-338 drops the non-finite rows of the patch first: redo this patch alone on the filtered rows
                Xb, yb = _np(X[b]), _np(y[b])
                ok = np.isfinite(Xb).all(axis=1) & np.isfinite(yb)
                if ok.sum() < 100:
                    res.append(None)
                    continue
                cb, rb = fit_patch_rows(ops._dev(np.ascontiguousarray(Xb[ok]))[None], ops._dev(np.ascontiguousarray(yb[ok]))[None], alpha)
                coeffs[b], r2[b] = cb[0], rb[0]
            elif n < 100:                                          # sindy:340-341: too few points
                res.append(None)
                continue
            res.append(coeffs[b])
            qual[b] = max(0.0, r2[b])                              # reg_quality = 1.0 without registration (sindy:290-293)
        return res, qual

    def discover_pde_for_patch(self, patch_sequence, skip_boundary=5, subsample=4, alpha=0.01, registration_method="none"):
        """sindy:272-366 for one patch location: (coeffs or None, quality)."""
        if registration_method != "none":
            raise NotImplementedError("patch registration (ECC / optical flow) is OpenCV pre-processing, not on the GPU path")
        seq = np.ascontiguousarray(np.asarray(patch_sequence, dtype=np.float64))
        if seq.ndim != 3 or seq.shape[1] != seq.shape[2]:
            raise ValueError("patch_sequence must be (T, h, h)")
        if seq.shape[0] < 3:
            return None, 0.0                                        # sindy:331-332: no interior frame
        keep = self.patch_size
        self.patch_size = seq.shape[1]
        try:
            res, qual = self._fit(seq, [(0, 0)], alpha, skip_boundary, subsample)
        finally:
            self.patch_size = keep
        return (res[0], float(qual[0])) if res[0] is not None else (None, 0.0)

    def discover_pde_patch_ensemble(self, alpha=0.01, min_patches=5, registration_method="none", max_patches=None):
        """sindy:368-490: every patch location of ``self.images`` (a list of equally sized float images), then the
        quality-weighted ensemble; returns (coeffs_ensemble, term_names, info) like the reference (None, None, {} when
        fewer than ``min_patches`` patches qualify)."""
        if registration_method != "none":
            raise NotImplementedError("patch registration (ECC / optical flow) is OpenCV pre-processing, not on the GPU path")
        U = np.ascontiguousarray(np.asarray(self.images, dtype=np.float64))
        origins = self.patch_origins(U.shape[1:])
        idx = list(range(len(origins)))
        if max_patches and len(origins) > max_patches:
            import random

            idx = random.sample(range(len(origins)), max_patches)       # sindy:412-415 (Python's global RNG, like the script)
        res, qual = self._fit(U, [origins[i] for i in idx], alpha, 5, 4)
        ok = [b for b in range(len(idx)) if res[b] is not None and qual[b] > -0.5]          # sindy:437
        if len(ok) < min_patches:
            return None, None, {}
        C_ = np.array([res[b] for b in ok])
        q = np.array([qual[b] for b in ok])
        weights = q / q.sum()                                                              # sindy:456-470
        ens = np.average(C_, axis=0, weights=weights)
        std = np.sqrt(np.average((C_ - ens) ** 2, axis=0, weights=weights))
        ens[std > np.median(std) * 2] = 0
        # the reference's metrics dict (sindy:484-490; no wall time here) plus the per-patch table it does not return
        return ens, list(TERM_NAMES), dict(n_patches=len(ok), avg_quality=float(q.mean()), quality_std=float(q.std()),
                                           coeffs_std=std, patch_coeffs=C_, patch_qualities=q,
                                           origins=[origins[idx[b]] for b in ok])
