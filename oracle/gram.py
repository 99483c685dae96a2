"""Oracle: sufficient statistics and STRidge-from-statistics (what kernels K1/K3 compute).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  All three reference STRidge
dialects depend on the rows only through {n, sum(theta), sum(y), sum(y^2), Theta^T Theta,
Theta^T y}; these functions restate each dialect on those statistics so that the GPU
path (which never materialises Theta) has a CPU twin.  tests/test_oracle_golden.py pins
them to the row-form oracles (and through those to the reference).

Statistics vector layout (shared with include/pdegram.h, ``PG_STATS_LEN(p)``):
    [0] n   [1] sum y   [2] sum y^2   [3 : 3+p] sum theta_j   [3+p : 3+2p] sum theta_j*y
    [3+2p : ] upper triangle of Theta^T Theta, row-major (i <= j)
"""

from __future__ import annotations

import numpy as np

DIALECT_KS, DIALECT_SKLEARN, DIALECT_BASIC = 0, 1, 2


def stats_len(p: int) -> int:
    return 3 + 2 * p + p * (p + 1) // 2


def pack_stats(X, y):
    """Rows -> statistics vector (float64)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    p = X.shape[1]
    G = X.T @ X
    iu = np.triu_indices(p)
    return np.concatenate([[float(len(y)), y.sum(), (y * y).sum()], X.sum(axis=0), X.T @ y, G[iu]])


def unpack_stats(s, p: int):
    s = np.asarray(s, dtype=np.float64)
    n, sy, syy = s[0], s[1], s[2]
    sx = s[3:3 + p].copy()
    b = s[3 + p:3 + 2 * p].copy()
    G = np.zeros((p, p))
    iu = np.triu_indices(p)
    G[iu] = s[3 + 2 * p:3 + 2 * p + p * (p + 1) // 2]
    G = G + np.triu(G, 1).T
    return n, sy, syy, sx, b, G


def rescale_stats(s, p: int, d):
    """Statistics of X / d (column scaling), e.g. the train-RMS scale of ks2d:1647-1655."""
    n, sy, syy, sx, b, G = unpack_stats(s, p)
    d = np.asarray(d, dtype=np.float64)
    G2 = G / np.outer(d, d)
    iu = np.triu_indices(p)
    return np.concatenate([[n, sy, syy], sx / d, b / d, G2[iu]])


def rms_scale(s, p: int, const_cols=()):
    """ks2d:1647-1652 from the train statistics: sqrt(G_jj / n) + 1e-12, '1' column -> 1."""
    n, _, _, _, _, G = unpack_stats(s, p)
    scale = np.sqrt(np.diag(G) / n) + 1e-12
    for j in const_cols:
        scale[j] = 1.0
    return scale


def _standardised_system(s, p, dialect, const_cols):
    """Centred/scaled normal equations  (C, r, scale)  with exact zeros for constant columns.

    ks2d:43-52 / patch:79-80: Xs = (X - mean) / std.  Xs^T Xs = (G - n mu mu^T) / (s s^T) and
    Xs^T y = (b - mu * sum y) / s  (centring y as sklearn's intercept does leaves Xs^T y
    unchanged because Xs is centred).  A constant column is exactly 0 after centring in the
    reference, so its row/column and rhs are forced to exact zero and its scale is 1.
    """
    n, sy, syy, sx, b, G = unpack_stats(s, p)
    mu = sx / n
    C = G - n * np.outer(mu, mu)
    var = np.diag(C) / n
    const = np.zeros(p, dtype=bool)
    for j in const_cols:
        const[j] = True
    if dialect == DIALECT_SKLEARN:
        eps = np.finfo(np.float64).eps
        const |= var <= n * eps * var + (n * mu * eps) ** 2
    const |= ~(var > 0)
    scale = np.where(const, 1.0, np.sqrt(np.where(var > 0, var, 1.0)))
    C = C / np.outer(scale, scale)
    r = (b - mu * sy) / scale
    C[const, :] = 0.0
    C[:, const] = 0.0
    r[const] = 0.0
    return C, r, scale


def _solve(A, rhs, dialect):
    if dialect == DIALECT_SKLEARN:
        from scipy import linalg

        return linalg.solve(A, rhs, assume_a="pos")
    return np.linalg.solve(A, rhs)


def stridge_from_stats(s, p: int, *, dialect: int, alpha: float, threshold: float, max_iter: int,
                       const_cols=(), signs=None):
    """STRidge on statistics.  ks2d:404-428 (DIALECT_KS), patch:78-98 (DIALECT_SKLEARN),
    basic:104-143 (DIALECT_BASIC).  ``signs`` (KS dialect): stridge_sign_constrained, ks2d:552-600 --
    wrong-signed coefficients are zeroed before the threshold mask and after every refit."""
    if dialect == DIALECT_BASIC:
        n, sy, syy, sx, b, G = unpack_stats(s, p)
        coef = np.ones(p)
        for _ in range(max_iter):
            coef = np.linalg.solve(G + alpha * np.eye(p), b)
            mask = np.abs(coef) < threshold
            coef[mask] = 0
            act = ~mask
            k = int(act.sum())
            if k == 0:
                break
            coef[act] = np.linalg.solve(G[np.ix_(act, act)] + alpha * np.eye(k), b[act])
        return coef

    C, r, scale = _standardised_system(s, p, dialect, const_cols)
    c = _solve(C + alpha * np.eye(p), r, dialect)

    def enforce(c):
        if signs is not None:
            for j, sg in enumerate(signs):
                if (sg == -1 and c[j] > 0) or (sg == 1 and c[j] < 0):
                    c[j] = 0.0

    for _ in range(max_iter):
        enforce(c)
        small = np.abs(c) < threshold
        if small.all():
            c = np.zeros(p)
            break
        big = ~small
        k = int(big.sum())
        cb = _solve(C[np.ix_(big, big)] + alpha * np.eye(k), r[big], dialect)
        c = np.zeros(p)
        c[big] = cb
        enforce(c)
    return c / (scale + 1e-12)


def metrics_from_stats(s, p: int, coef):
    """r2 (ks2d:35-40 incl. the 1e-18 guard) and rmse (ks2d:29-32) of ``X @ coef`` from the
    statistics of the evaluation rows: ss_res = yy - 2 c.b + c.G.c.  Suffers cancellation
    when the fit is exact to ~1e-8 relative (documented in DESIGN.md)."""
    n, sy, syy, sx, b, G = unpack_stats(s, p)
    coef = np.asarray(coef, dtype=np.float64)
    ss_res = syy - 2.0 * coef @ b + coef @ G @ coef
    ss_res = max(ss_res, 0.0)
    ss_tot = syy - sy * sy / n
    return float(1.0 - ss_res / (ss_tot + 1e-18)), float(np.sqrt(ss_res / n))


def ks_fit_from_stats(s_train, s_test, p: int, *, alphas, thresholds, const_cols=(), max_iter: int = 25):
    """ks2d:1638-1743 on statistics: RMS scale from the train Gram diagonal, STRidge per
    (alpha, threshold), unscale, test r2/rmse, best key (r2, -n_active, -rmse), first max wins."""
    scale = rms_scale(s_train, p, const_cols)
    st = rescale_stats(s_train, p, scale)
    best = None
    table = []
    for a in alphas:
        for thr in thresholds:
            c = stridge_from_stats(st, p, dialect=DIALECT_KS, alpha=a, threshold=thr, max_iter=max_iter,
                                   const_cols=const_cols) / scale
            r2, err = metrics_from_stats(s_test, p, c)
            na = int(np.sum(np.abs(c) > 0))
            table.append((a, thr, r2, err, na))
            key = (r2, -na, -err)
            if best is None or key > best["key"]:
                best = dict(key=key, alpha=a, threshold=thr, coeffs=c, r2_test=r2, rmse_test=err, n_active=na)
    best["table"] = table
    best["scale"] = scale
    return best
