"""Oracle: the KS-2D dialect (periodic FD, dict libraries, blockwise means, STRidge).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  NumPy restatement of
scripts/ks2d_stridge_benchmark.py ("ks2d").  Layout is ``U[t, x, y]`` with "x"
on axis 0 of a frame (ks2d:70-73, ks2d:1295).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

GRID_ALPHAS = (1e-6, 1e-5, 1e-4, 1e-3, 1e-2)            # ks2d:1721
GRID_THRESHOLDS = (1e-10, 1e-9, 1e-8, 1e-7, 1e-6, 1e-5)  # ks2d:1722


# --------------------------------------------------------------------------- metrics
def rmse(y_true, y_pred) -> float:
    """ks2d:29-32."""
    d = np.asarray(y_true) - np.asarray(y_pred)
    return float(np.sqrt(np.mean(d ** 2)))


def r2_score(y_true, y_pred) -> float:
    """ks2d:35-40 (note the 1e-18 guard in the denominator)."""
    y_true = np.asarray(y_true)
    y_pred = np.asarray(y_pred)
    ss_res = float(np.sum((y_true - y_pred) ** 2))
    ss_tot = float(np.sum((y_true - float(np.mean(y_true))) ** 2))
    return float(1.0 - ss_res / (ss_tot + 1e-18))


# --------------------------------------------------------------------------- stencils
def laplacian(f, dx: float, dy: float):
    """ks2d:63-67.  Periodic 5-point Laplacian on the last two axes of ``f``.

    The reference is called per 2-D frame; np.roll on axes (-2, -1) of a stack is
    the same arithmetic applied frame by frame.
    """
    f = np.asarray(f)
    ax0, ax1 = f.ndim - 2, f.ndim - 1
    d0 = (np.roll(f, -1, axis=ax0) - 2 * f + np.roll(f, 1, axis=ax0)) / (dx ** 2)
    d1 = (np.roll(f, -1, axis=ax1) - 2 * f + np.roll(f, 1, axis=ax1)) / (dy ** 2)
    return d0 + d1


def gradients(f, dx: float, dy: float):
    """ks2d:70-73.  Periodic central differences; gx along frame axis 0, gy along axis 1."""
    f = np.asarray(f)
    ax0, ax1 = f.ndim - 2, f.ndim - 1
    gx = (np.roll(f, -1, axis=ax0) - np.roll(f, 1, axis=ax0)) / (2 * dx)
    gy = (np.roll(f, -1, axis=ax1) - np.roll(f, 1, axis=ax1)) / (2 * dy)
    return gx, gy


def ks_rhs(u, dx: float, dy: float):
    """ks2d:118-122."""
    lap = laplacian(u, dx, dy)
    bih = laplacian(lap, dx, dy)
    ux, uy = gradients(u, dx, dy)
    return -lap - bih - 0.5 * (ux ** 2 + uy ** 2)


# --------------------------------------------------------------------------- input generation
@dataclass(frozen=True)
class SimConfig:
    """ks2d:751-760."""

    Lx: float = 50.0
    Ly: float = 50.0
    Nx: int = 100
    Ny: int = 100
    dt: float = 1e-3
    n_seconds: float = 2.0
    save_every: int = 1
    seed: int = 42


def simulate(cfg: SimConfig):
    """ks2d:763-782.  Explicit-Euler KS; used only to MAKE config-1/2 inputs."""
    dx = cfg.Lx / cfg.Nx
    dy = cfg.Ly / cfg.Ny
    total_steps = int(cfg.n_seconds / cfg.dt)
    n_frames = total_steps // cfg.save_every
    DT = cfg.dt * cfg.save_every
    rng = np.random.default_rng(cfg.seed)
    u = rng.uniform(-0.1, 0.1, size=(cfg.Nx, cfg.Ny)).astype(np.float64)
    U = np.zeros((n_frames, cfg.Nx, cfg.Ny), dtype=np.float64)
    k = 0
    for step in range(total_steps):
        u = np.nan_to_num(u + cfg.dt * ks_rhs(u, dx, dy))
        if step % cfg.save_every == 0:
            U[k] = u
            k += 1
    return U, dx, dy, DT


def add_noise(U, noise_rel: float, seed: int = 999):
    """ks2d:1409 + ks2d:840-845 (perturbation N2_noise): sigma = noise_rel * std(U)."""
    U = np.asarray(U, dtype=np.float64).copy()
    if noise_rel <= 0:
        return U
    rng = np.random.default_rng(int(seed))
    sigma = float(noise_rel) * float(np.std(U))
    return U + rng.normal(0.0, sigma, size=U.shape)


# --------------------------------------------------------------------------- libraries
RICH_NAMES = ["1", "u", "u^2", "u_x", "u_y", "∇²u", "∇⁴u", "|∇u|²", "u·∇²u"]  # ks2d:1048-1059
TRUE_NAMES = ["∇²u", "∇⁴u", "|∇u|²"]                                         # ks2d:1095-1099


def build_dictionary(U_mid, dx: float, dy: float, *, deriv: str = "finite", spectral_cutoff: float = 1.0):
    """ks2d:1017-1060.  p=9 rich dictionary; dict insertion order is the column order."""
    if deriv != "finite":
        raise NotImplementedError("oracle restates the finite-difference path only")
    U_mid = np.asarray(U_mid)
    ux, uy = gradients(U_mid, dx, dy)
    lap = laplacian(U_mid, dx, dy)
    bih = laplacian(lap, dx, dy)
    terms = {
        "1": np.ones_like(U_mid),
        "u": U_mid,
        "u^2": U_mid ** 2,
        "u_x": ux,
        "u_y": uy,
        "∇²u": lap,
        "∇⁴u": bih,
        "|∇u|²": ux ** 2 + uy ** 2,
        "u·∇²u": U_mid * lap,
    }
    return list(terms.keys()), terms


def build_dictionary_true(U_frames, dx: float, dy: float, *, deriv: str = "finite",
                          spectral_cutoff: float = 1.0, include_advection: bool = False):
    """ks2d:1063-1104.  p=3 (or 5 with advection) dictionary matching the KS RHS."""
    if deriv != "finite":
        raise NotImplementedError("oracle restates the finite-difference path only")
    U_frames = np.asarray(U_frames)
    ux, uy = gradients(U_frames, dx, dy)
    lap = laplacian(U_frames, dx, dy)
    bih = laplacian(lap, dx, dy)
    terms = {"∇²u": lap, "∇⁴u": bih, "|∇u|²": ux ** 2 + uy ** 2}
    if include_advection:
        terms["u_x"] = ux
        terms["u_y"] = uy
    return list(terms.keys()), terms


# --------------------------------------------------------------------------- blockwise means
def _block_mean(a, bt: int, bx: int, by: int):
    """Mean over (bt,bx,by) blocks, ragged trailing blocks kept (ks2d:380-389)."""
    T, nx, ny = a.shape
    it = np.arange(0, T, bt)
    ix = np.arange(0, nx, bx)
    iy = np.arange(0, ny, by)
    s = np.add.reduceat(a, it, axis=0)
    s = np.add.reduceat(s, ix, axis=1)
    s = np.add.reduceat(s, iy, axis=2)
    ct = np.minimum(it + bt, T) - it
    cx = np.minimum(ix + bx, nx) - ix
    cy = np.minimum(iy + by, ny) - iy
    cnt = ct[:, None, None] * cx[None, :, None] * cy[None, None, :]
    return s / cnt


def build_blockwise_dataset(Ut, terms, names, *, block_t: int, block_x: int, block_y: int):
    """ks2d:358-401.  Row order: t-block major, then x-block, then y-block; rows with a
    non-finite mean are dropped (ks2d:394-395)."""
    Ut = np.asarray(Ut)
    if Ut.ndim != 3:
        raise ValueError("Ut must be (T, Nx, Ny)")
    bt, bx, by = int(block_t), int(block_x), int(block_y)
    if bt <= 0 or bx <= 0 or by <= 0:
        raise ValueError("block sizes must be > 0")
    y = _block_mean(Ut, bt, bx, by).reshape(-1)
    X = np.stack([_block_mean(np.asarray(terms[n]), bt, bx, by).reshape(-1) for n in names], axis=1)
    keep = np.isfinite(y) & np.isfinite(X).all(axis=1)
    if not keep.any():
        return np.zeros((0, len(names))), np.zeros((0,))
    return X[keep], y[keep]


def build_blockwise_dataset_loops(Ut, terms, names, *, block_t: int, block_x: int, block_y: int):
    """Literal triple-loop form of ks2d:380-401 (small cases; pins the vectorised one)."""
    T, nx, ny = Ut.shape
    rows, ys = [], []
    for t0 in range(0, T, block_t):
        for x0 in range(0, nx, block_x):
            for y0 in range(0, ny, block_y):
                sl = (slice(t0, min(T, t0 + block_t)), slice(x0, min(nx, x0 + block_x)),
                      slice(y0, min(ny, y0 + block_y)))
                yb = float(np.mean(Ut[sl]))
                xb = np.array([float(np.mean(terms[n][sl])) for n in names])
                if np.isfinite(yb) and np.isfinite(xb).all():
                    rows.append(xb)
                    ys.append(yb)
    if not rows:
        return np.zeros((0, len(names))), np.zeros((0,))
    return np.stack(rows), np.asarray(ys)


# --------------------------------------------------------------------------- optional denoising prologue
def gaussian_smooth_periodic_2d(frame, sigma_px: float):
    """ks2d:125-142: multiply the 2-D FFT by exp(-sigma^2 (kx^2 + ky^2) / 2)."""
    sigma_px = float(sigma_px)
    if sigma_px <= 0:
        return frame.astype(np.float64, copy=True)
    nx, ny = frame.shape
    kx = 2.0 * np.pi * np.fft.fftfreq(nx)
    ky = 2.0 * np.pi * np.fft.fftfreq(ny)
    KX, KY = np.meshgrid(kx, ky, indexing="ij")
    H = np.exp(-0.5 * (sigma_px ** 2) * (KX ** 2 + KY ** 2))
    return np.fft.ifft2(np.fft.fft2(frame.astype(np.float64, copy=False)) * H).real


def time_smooth_moving_average(U, window: int):
    """ks2d:145-161: reflect padding, cumulative sum, difference / window."""
    window = int(window)
    if window <= 1:
        return U.astype(np.float64, copy=True)
    if window % 2 == 0:
        raise ValueError("time smoothing window must be odd")
    pad = window // 2
    U_pad = np.pad(U.astype(np.float64, copy=False), ((pad, pad), (0, 0), (0, 0)), mode="reflect")
    cs = np.concatenate([np.zeros_like(U_pad[:1]), np.cumsum(U_pad, axis=0)], axis=0)
    return (cs[window:] - cs[:-window]) / float(window)


# --------------------------------------------------------------------------- STRidge
def standardize_fit(X):
    """ks2d:43-48: column mean and population std; zero std -> 1."""
    mean = np.mean(X, axis=0)
    scale = np.std(X, axis=0)
    return mean, np.where(scale > 0, scale, 1.0)


def standardize_transform(X, mean, scale):
    """ks2d:51-52."""
    return (X - mean) / scale


def ridge_fit(X, y, alpha: float):
    """ks2d:55-60: solve (X^T X + alpha I) b = X^T y, no intercept."""
    G = X.T @ X
    return np.linalg.solve(G + alpha * np.eye(G.shape[0]), X.T @ y)


def stridge(X, y, *, alpha: float = 1e-3, threshold: float = 1e-6, max_iter: int = 25):
    """ks2d:404-428.  X is centred and scaled, y is NOT centred; strict '<' threshold;
    always ``max_iter`` refits unless every coefficient is small; returns c/(scale+1e-12)."""
    mean, scale = standardize_fit(X)
    Xs = standardize_transform(X, mean, scale)
    c = ridge_fit(Xs, y, alpha).copy()
    for _ in range(max_iter):
        small = np.abs(c) < threshold
        if small.all():
            c[:] = 0.0
            break
        big = ~small
        cb = ridge_fit(Xs[:, big], y, alpha)
        c = np.zeros_like(c)
        c[big] = cb
    return c / (scale + 1e-12)


def _enforce_signs(c, signs):
    """ks2d:577-582 / 594-598: a coefficient whose sign contradicts signs[j] (-1 / +1) becomes 0."""
    for j, sg in enumerate(signs):
        if (sg == -1 and c[j] > 0) or (sg == 1 and c[j] < 0):
            c[j] = 0.0


def stridge_sign_constrained(X, y, *, alpha: float = 1e-3, threshold: float = 1e-6, max_iter: int = 25, signs=None):
    """ks2d:552-600.  STRidge (same standardisation and refits as ks2d:404-428) where, inside every
    iteration, wrong-signed coefficients are zeroed BEFORE the threshold mask is formed and again after
    the refit; the initial full ridge fit is not sign-filtered when max_iter == 0."""
    mean, scale = standardize_fit(X)
    Xs = standardize_transform(X, mean, scale)
    p = X.shape[1]
    signs = [0] * p if signs is None else list(signs)
    c = ridge_fit(Xs, y, alpha).copy()
    for _ in range(max_iter):
        _enforce_signs(c, signs)
        small = np.abs(c) < threshold
        if small.all():
            c[:] = 0.0
            break
        big = ~small
        cb = ridge_fit(Xs[:, big], y, alpha)
        c = np.zeros_like(c)
        c[big] = cb
        _enforce_signs(c, signs)
    return c / (scale + 1e-12)


def ensemble_stridge(X, y, *, alpha: float = 1e-3, threshold: float = 1e-6, max_iter: int = 25, n_bootstrap: int = 50,
                     subsample_frac: float = 0.7, seed: int = 0, return_all: bool = False):
    """ks2d:603-642 (use_huber=False): bootstrap resamples with replacement, stridge on each, median / std."""
    rng = np.random.default_rng(seed)
    n = len(y)
    n_sub = max(int(n * subsample_frac), 1)
    allc = []
    for _ in range(n_bootstrap):
        idx = rng.choice(n, size=n_sub, replace=True)
        allc.append(stridge(X[idx], y[idx], alpha=alpha, threshold=threshold, max_iter=max_iter))
    allc = np.stack(allc, axis=0)
    if return_all:
        return allc
    return np.median(allc, axis=0), np.std(allc, axis=0)


def rollout_errors(U, dx, dy, DT, names, coeffs, n_steps: int = 50):
    """ks2d:1804-1838: u_hat <- u_hat + DT * rhs_from_coeffs(u_hat) from U[0]; rmse(U[k+1], u_hat) per step.
    rhs: out = zeros; for (name, c): skip |c| < 1e-12; out += c * value(name)."""
    def rhs(u):
        ux, uy = gradients(u, dx, dy)
        lap = laplacian(u, dx, dy)
        bih = laplacian(lap, dx, dy)
        vals = {"1": 1.0, "u": u, "u^2": u ** 2, "u_x": ux, "u_y": uy, "∇²u": lap, "∇⁴u": bih,
                "|∇u|²": ux ** 2 + uy ** 2, "u·∇²u": u * lap}
        out = np.zeros_like(u, dtype=np.float64)
        for name, c in zip(names, coeffs):
            if abs(c) < 1e-12:
                continue
            v = vals[name]
            out += c * (v if isinstance(v, np.ndarray) else float(v))
        return out

    n = int(min(n_steps, U.shape[0] - 1))
    u_hat = U[0].copy()
    errs = []
    for k in range(n):
        u_hat = u_hat + DT * rhs(u_hat)
        errs.append(rmse(U[k + 1].ravel(), u_hat.ravel()))
    return np.asarray(errs)


# --------------------------------------------------------------------------- main() hot path
def make_dataset(U, dx, dy, DT, *, method: str = "pointwise", dictionary: str = "true",
                 include_advection: bool = False, n_sample: int = 50_000,
                 block=(3, 8, 8), rng=None):
    """ks2d:1508-1550 (blockwise) and ks2d:1551-1636 (pointwise).  ``rng`` is the
    ``default_rng(0)`` stream created at ks2d:1470 and is advanced exactly as there."""
    if rng is None:
        rng = np.random.default_rng(0)
    U = np.asarray(U, dtype=np.float64)
    U_frames = U[:-1]
    Ut = (U[1:] - U[:-1]) / DT
    if dictionary == "true":
        names, terms = build_dictionary_true(U_frames, dx, dy, include_advection=include_advection)
    else:
        names, terms = build_dictionary(U_frames, dx, dy)
    if method == "blockwise":
        X, y = build_blockwise_dataset(Ut, terms, names, block_t=block[0], block_x=block[1], block_y=block[2])
    elif method == "pointwise":
        n_total = Ut.size
        flat_idx = rng.choice(n_total, size=int(min(n_sample, n_total)), replace=False)
        y = Ut.reshape(-1)[flat_idx]
        X = np.column_stack([terms[n].reshape(-1)[flat_idx] for n in names])
        ok = np.isfinite(X).all(axis=1) & np.isfinite(y)
        X, y = X[ok], y[ok]
    else:
        raise ValueError(method)
    return names, X, y, rng


def split_and_scale(names, X, y, rng):
    """ks2d:1638-1655: 70/30 permutation split; train-RMS column scale (+1e-12), '1' -> 1."""
    perm = rng.permutation(len(y))
    split = int(0.7 * len(y))
    tr, te = perm[:split], perm[split:]
    scale = np.sqrt(np.mean(X[tr] ** 2, axis=0)) + 1e-12
    for j, n in enumerate(names):
        if n == "1":
            scale[j] = 1.0
    return tr, te, scale


def fit(names, X, y, rng, *, alpha: float = 1e-6, threshold: float = 1e-10, grid_search: bool = False):
    """ks2d:1638-1779: split, RMS-scale, STRidge (or the 5x6 sweep), test metrics."""
    tr, te, scale = split_and_scale(names, X, y, rng)
    X_tr, y_tr, X_te, y_te = X[tr], y[tr], X[te], y[te]
    X_tr_s = X_tr / scale

    def one(a, thr):
        c = stridge(X_tr_s, y_tr, alpha=a, threshold=thr, max_iter=25) / scale
        pred = X_te @ c
        return c, r2_score(y_te, pred), rmse(y_te, pred), int(np.sum(np.abs(c) > 0))

    if grid_search:
        best = None
        table = []
        for a in GRID_ALPHAS:
            for thr in GRID_THRESHOLDS:
                c, r2, err, na = one(a, thr)
                table.append((a, thr, r2, err, na))
                key = (r2, -na, -err)
                if best is None or key > best["key"]:  # ks2d:1732
                    best = dict(key=key, alpha=a, threshold=thr, coeffs=c, r2_test=r2, rmse_test=err, n_active=na)
        best["table"] = table
    else:
        c, r2, err, na = one(alpha, threshold)
        best = dict(alpha=alpha, threshold=threshold, coeffs=c, r2_test=r2, rmse_test=err, n_active=na)
    best.update(names=names, train_idx=tr, test_idx=te, scale=scale,
                r2_train=r2_score(y_tr, X_tr @ best["coeffs"]), rmse_train=rmse(y_tr, X_tr @ best["coeffs"]))
    return best


def run_config(U, dx, dy, DT, **kw):
    """Dataset + fit in one call (the reference's main() hot path for one config)."""
    fit_kw = {k: kw.pop(k) for k in ("alpha", "threshold", "grid_search") if k in kw}
    names, X, y, rng = make_dataset(U, dx, dy, DT, **kw)
    out = fit(names, X, y, rng, **fit_kw)
    out.update(X_shape=X.shape, X=X, y=y)
    return out
