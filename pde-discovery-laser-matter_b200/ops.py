"""Device-level operators: one Python function per C-ABI entry point of libpdegram.so.

Inputs may be NumPy arrays (copied to the current CUDA device) or torch CUDA tensors (used
in place); outputs are torch CUDA tensors.  All calls are asynchronous on torch's current
stream.  The dialect modules (ks2d, basic_usage, patch) build the reference signatures on
top of these.
"""

from __future__ import annotations

import numpy as np

from . import _lib as L


def _dev(a, dtype=None):
    """NumPy array / torch tensor -> contiguous CUDA tensor of `dtype` (torch dtype)."""
    torch = L.torch_cuda()
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        t = a
        if not t.is_cuda:
            t = t.cuda()
    else:
        from . import _xfer

        t = _xfer.to_device(np.asarray(a))     # large arrays: pinned double-buffered staging
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def row_space(shape, dialect):
    """(row frames, R0, R1) of a (T, A0, A1) stack: KS every point of U[:-1]; basic_usage U[:-1, 2:-2, 2:-2];
    analyze_results U[:-2, :-2, :-2] (central time difference, slice-aligned stencils)."""
    T, A0, A1 = (int(x) for x in shape)
    if dialect == L.FD_KS_PERIODIC:
        return max(T - 1, 0), A0, A1
    if dialect == L.FD_BASIC_TRIM:
        return max(T - 1, 0), A0 - 4, A1 - 4
    if dialect == L.FD_SLICE_CENTRAL:
        return max(T - 2, 0), A0 - 2, A1 - 2
    raise ValueError(f"unknown finite-difference dialect {dialect}")


def field(U):
    """A (T, A0, A1) float64 stack on the device."""
    torch = L.torch_cuda()
    t = _dev(U, torch.float64)
    if t.ndim != 3:
        raise ValueError("field must be (T, A0, A1)")
    return t


def fd_lib_gram(U, d0, d1, dt, *, dialect, library, block=(1, 1, 1), fold_of_row=None, fold_of_frame=None,
                n_folds=1, variant=L.VARIANT_AUTO, return_nonfinite=False, trailing_block_means=None, halo=None, Uy=None):
    """K1: field -> statistics [n_folds][S(p)] without materialising Theta (pg_fd_lib_gram).

    ``return_nonfinite``: also return the call's four int64 counters (device tensor): [0] block rows dropped because a
    mean was not finite, [1] rows with a fold id outside [0, n_folds) (a caller error: the statistics are then NaN),
    [2] internal, [3] halo waits that timed out (statistics NaN), [4..7] SM cycle counter / globaltimer ns at the start
    and end of the tiled blockwise kernel's CTA 0 (effective SM clock of the launch).

    ``Uy``: a second stack of U's shape that the time derivative is taken of (the library still comes from U):
    pg_fd_lib_gram_two, the ks2d script's ``--denoise-space-on features``.  (bt, 8, 8) blocks go through the tiled kernel
    (the block mean of u_t telescopes to block sums of Uy's frames k bt), pointwise rows through the generic kernel.

    ``halo`` = (flag_ptr, epoch) from ``slabs.PeerComm.pull_halo``: U[-1] is still being filled by a copy engine; the
    kernel starts at once and reads that frame only after the flag has reached ``epoch`` (pg_fd_lib_gram_halo).

    ``trailing_block_means`` ([A0/8][A1/8], see ``frame_block_means``): the (8, 8) block means of the frame that
    follows U[-2]; U[-1] is then a placeholder whose values are ignored (pg_fd_lib_gram_tail: time slabs whose
    trailing frame lives on another GPU)."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, A0, A1 = U.shape
    p = L.LIB_WIDTH[library]
    bt, b0, b1 = (int(b) for b in block)
    fr = _dev(fold_of_row, torch.uint8)
    ff = _dev(fold_of_frame, torch.int32)
    Tr, R0, R1 = row_space(U.shape, dialect)
    if fr is not None:
        nrows = -(-Tr // bt) * -(-R0 // b0) * -(-R1 // b1)
        if fr.numel() != nrows:
            raise ValueError(f"fold_of_row has {fr.numel()} entries, the block grid has {nrows} rows")
    if ff is not None and ff.numel() != Tr:
        raise ValueError(f"fold_of_frame must have one entry per row frame = {Tr} entries")
    stats = torch.empty((n_folds, L.stats_len(p)), dtype=torch.float64, device=U.device)
    bad = torch.zeros(8, dtype=torch.int64, device=U.device) if return_nonfinite else None
    if Uy is not None:
        Uy = field(Uy)
        if Uy.shape != U.shape:
            raise ValueError("Uy must have the shape of U")
        L.check(lib.pg_fd_lib_gram_two(L.ptr(U), L.ptr(Uy), T, A0, A1, float(d0), float(d1), float(dt), dialect, library, bt, b0,
                                       b1, L.ptr(fr), L.ptr(ff), n_folds, L.ptr(stats), L.ptr(bad), variant, L.stream_ptr()))
    elif halo is not None:
        L.check(lib.pg_fd_lib_gram_halo(L.ptr(U), T, A0, A1, float(d0), float(d1), float(dt), dialect, library, bt, b0,
                                        b1, L.ptr(fr), L.ptr(ff), n_folds, int(halo[0]), int(halo[1]), L.ptr(stats),
                                        L.ptr(bad), variant, L.stream_ptr()))
    elif trailing_block_means is not None:
        tm = _dev(trailing_block_means, torch.float64)
        if tm.numel() != (A0 // 8) * (A1 // 8):
            raise ValueError(f"trailing_block_means must hold (A0/8) x (A1/8) = {(A0 // 8) * (A1 // 8)} block means")
        L.check(lib.pg_fd_lib_gram_tail(L.ptr(U), T, A0, A1, float(d0), float(d1), float(dt), dialect, library, bt, b0,
                                        b1, L.ptr(fr), L.ptr(ff), n_folds, L.ptr(tm), L.ptr(stats), L.ptr(bad), variant,
                                        L.stream_ptr()))
    else:
        L.check(lib.pg_fd_lib_gram(L.ptr(U), T, A0, A1, float(d0), float(d1), float(dt), dialect, library, bt, b0, b1,
                                   L.ptr(fr), L.ptr(ff), n_folds, L.ptr(stats), L.ptr(bad), variant, L.stream_ptr()))
    return (stats, bad) if return_nonfinite else stats


def frame_block_means(frame, out=None):
    """(8, 8) block means [A0/8][A1/8] of one (A0, A1) frame (pg_block_means): what a rank publishes of its first
    frame instead of the frame itself (``fd_lib_gram(..., trailing_block_means=...)``)."""
    torch = L.torch_cuda()
    lib = L.load()
    f = _dev(frame, torch.float64)
    A0, A1 = f.shape
    if A0 % 8 or A1 % 8:
        raise ValueError("frame_block_means needs extents that are multiples of 8")
    if out is None:
        out = torch.empty((A0 // 8, A1 // 8), dtype=torch.float64, device=f.device)
    L.check(lib.pg_block_means(L.ptr(f), 1, 1, A0, A1, 1, 8, 8, L.ptr(out), L.stream_ptr()))
    return out


def fd_terms(U, d0, d1, dt, *, dialect, library):
    """Materialised term stacks (pg_fd_terms): KS -> [p][T][A0][A1]; BASIC -> [5][T-1][A0-4][A1-4]."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, A0, A1 = U.shape
    if dialect == L.FD_BASIC_TRIM:
        shape = (5, max(T - 1, 0), max(A0 - 4, 0), max(A1 - 4, 0))
    else:
        shape = (L.LIB_WIDTH[library], T, A0, A1)
    out = torch.empty(shape, dtype=torch.float64, device=U.device)
    L.check(lib.pg_fd_terms(L.ptr(U), T, A0, A1, float(d0), float(d1), float(dt), dialect, library, L.ptr(out),
                            L.stream_ptr()))
    return out


def fd_gather_rows(U, d0, d1, dt, flat_idx, *, dialect, library, Uy=None):
    """K1c: sampled pointwise rows (pg_fd_gather_rows) -> (X [n][p], y [n]).  ``Uy``: the stack u_t is taken of, when
    it is not U (pg_fd_gather_rows_two)."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, A0, A1 = U.shape
    idx = _dev(flat_idx, torch.int64)
    n, p = idx.numel(), L.LIB_WIDTH[library]
    X = torch.empty((n, p), dtype=torch.float64, device=U.device)
    y = torch.empty((n,), dtype=torch.float64, device=U.device)
    if Uy is not None:
        Uy = field(Uy)
        if Uy.shape != U.shape:
            raise ValueError("Uy must have the shape of U")
        L.check(lib.pg_fd_gather_rows_two(L.ptr(U), L.ptr(Uy), T, A0, A1, float(d0), float(d1), float(dt), dialect, library,
                                          L.ptr(idx), n, L.ptr(X), L.ptr(y), L.stream_ptr()))
    else:
        L.check(lib.pg_fd_gather_rows(L.ptr(U), T, A0, A1, float(d0), float(d1), float(dt), dialect, library, L.ptr(idx), n,
                                      L.ptr(X), L.ptr(y), L.stream_ptr()))
    return X, y


def block_means(stack, block):
    """pg_block_means: [k][T][A0][A1] -> [nrows][k] in reference row order."""
    torch = L.torch_cuda()
    lib = L.load()
    s = _dev(stack, torch.float64)
    k, T, A0, A1 = s.shape
    bt, b0, b1 = (int(b) for b in block)
    if bt <= 0 or b0 <= 0 or b1 <= 0:
        raise ValueError("block sizes must be > 0")
    nrows = -(-T // bt) * -(-A0 // b0) * -(-A1 // b1)
    out = torch.empty((nrows, k), dtype=torch.float64, device=s.device)
    L.check(lib.pg_block_means(L.ptr(s), k, T, A0, A1, bt, b0, b1, L.ptr(out), L.stream_ptr()))
    return out


def rows_gram(X, y, *, fold_of_row=None, n_folds=1, shift=None, want_minmax=False):
    """pg_rows_gram: X [n][p] or [B][n][p], y [n] or [B][n] -> stats [B][n_folds][S] (+ colminmax)."""
    torch = L.torch_cuda()
    lib = L.load()
    X = _dev(X, torch.float64)
    y = _dev(y, torch.float64)
    if X.ndim == 2:
        X, y = X[None], y[None]
    B, n, p = X.shape
    if y.shape != (B, n):
        raise ValueError("y must match the rows of X")
    fr = _dev(fold_of_row, torch.uint8)
    sh = _dev(shift, torch.float64)
    if sh is not None:
        sh = sh.reshape(B, p).contiguous()
    stats = torch.empty((B, n_folds, L.stats_len(p)), dtype=torch.float64, device=X.device)
    mm = torch.empty((B, n_folds, 2, p), dtype=torch.float64, device=X.device) if want_minmax else None
    L.check(lib.pg_rows_gram(L.ptr(X), L.ptr(y), B, n, p, p, L.ptr(fr), n_folds, L.ptr(sh), L.ptr(stats), L.ptr(mm),
                             L.stream_ptr()))
    return (stats, mm) if want_minmax else stats


def rows_gram_weighted(X, y, weights, *, shift=None, want_minmax=False):
    """pg_rows_gram_weighted: statistics of B resamples of one row set.  X [n][p], y [n], weights [B][n]
    (multiplicities) -> stats [B][S] (+ colminmax [B][2][p] over the rows each resample drew)."""
    torch = L.torch_cuda()
    lib = L.load()
    X = _dev(X, torch.float64)
    y = _dev(y, torch.float64)
    n, p = X.shape
    w = _dev(weights).reshape(-1, n)
    if w.dtype != torch.uint16:
        if int(w.max()) > 65535:
            raise ValueError("row multiplicities above 65535 are not supported")
        w = w.to(torch.uint16)
    w = w.contiguous()
    B = w.shape[0]
    sh = None if shift is None else _dev(shift, torch.float64).reshape(p).contiguous()
    stats = torch.empty((B, L.stats_len(p)), dtype=torch.float64, device=X.device)
    mm = torch.empty((B, 2, p), dtype=torch.float64, device=X.device) if want_minmax else None
    L.check(lib.pg_rows_gram_weighted(L.ptr(X), L.ptr(y), n, p, p, L.ptr(w), B, L.ptr(sh), L.ptr(stats), L.ptr(mm),
                                      L.stream_ptr()))
    return (stats, mm) if want_minmax else stats


def poly_rows(U, pts, W6, rt, rs, *, library):
    """K2: pg_poly_rows.  U float32 or float64 (T,H,W); pts [n][3] (t,y,x) -> (X [n][p], y [n])."""
    torch = L.torch_cuda()
    lib = L.load()
    if isinstance(U, np.ndarray):
        if U.dtype not in (np.float32, np.float64):
            U = U.astype(np.float64)
        U = _dev(U)
    if U.dtype not in (torch.float32, torch.float64):
        U = U.to(torch.float64)
    U = U.contiguous()
    T, H, W = U.shape
    pts = _dev(pts, torch.int32).reshape(-1, 3).contiguous()
    W6 = _dev(W6, torch.float64)
    nnb = (2 * rt + 1) * (2 * rs + 1) ** 2
    if tuple(W6.shape) != (6, nnb):
        raise ValueError(f"W6 must be (6, {nnb})")
    n, p = pts.shape[0], L.LIB_WIDTH[library]
    X = torch.empty((n, p), dtype=torch.float64, device=U.device)
    y = torch.empty((n,), dtype=torch.float64, device=U.device)
    L.check(lib.pg_poly_rows(L.ptr(U), 0 if U.dtype == torch.float32 else 1, T, H, W, L.ptr(pts), n, L.ptr(W6), rt, rs,
                             library, L.ptr(X), L.ptr(y), L.stream_ptr()))
    return X, y


def stridge_batched(stats, p, *, dialect, alphas, thresholds, max_iter, flags=0, const_cols=(), colminmax=None,
                    shift=None, eval_stats=None, signs=None):
    """K3: pg_stridge_batched.  stats [B][S] -> dict(coef [B][na][nt][p], metrics, best)."""
    torch = L.torch_cuda()
    lib = L.load()
    stats = _dev(stats, torch.float64).reshape(-1, L.stats_len(p)).contiguous()
    B = stats.shape[0]
    def vec(v):  # sweep grids may already live on the device (no per-call copy)
        if isinstance(v, torch.Tensor):
            return _dev(v, torch.float64).reshape(-1)
        return _dev(np.atleast_1d(np.asarray(v, dtype=np.float64)))

    al, th = vec(alphas), vec(thresholds)
    na, nt = al.numel(), th.numel()
    cm = None
    if isinstance(const_cols, torch.Tensor):
        cm = _dev(const_cols, torch.uint8)
    elif len(tuple(const_cols)):
        m = np.zeros(p, dtype=np.uint8)
        m[list(const_cols)] = 1
        cm = _dev(m)
    sg = None
    if signs is not None:
        sg_h = np.asarray(list(signs), dtype=np.int64)
        if sg_h.shape != (p,) or not np.isin(sg_h, (-1, 0, 1)).all():
            raise ValueError("signs must hold p entries from {-1, 0, +1}")
        sg = _dev(sg_h.astype(np.int8))
    mm = None if colminmax is None else _dev(colminmax, torch.float64).reshape(B, 2, p).contiguous()
    sh = None if shift is None else _dev(shift, torch.float64).reshape(B, p).contiguous()
    ev = None if eval_stats is None else _dev(eval_stats, torch.float64).reshape(B, L.stats_len(p)).contiguous()
    coef = torch.empty((B, na, nt, p), dtype=torch.float64, device=stats.device)
    metrics = torch.empty((B, na, nt, 2), dtype=torch.float64, device=stats.device) if ev is not None else None
    best = torch.empty((B,), dtype=torch.int32, device=stats.device) if ev is not None else None
    relres = torch.empty((B, na, nt), dtype=torch.float64, device=stats.device) if ev is not None else None
    L.check(lib.pg_stridge_batched(L.ptr(stats), B, p, dialect, flags, L.ptr(al), na, L.ptr(th), nt, int(max_iter),
                                   L.ptr(cm), L.ptr(sg), L.ptr(mm), L.ptr(sh), L.ptr(ev), L.ptr(coef), L.ptr(metrics), L.ptr(best),
                                   L.ptr(relres), L.stream_ptr()))
    return dict(coef=coef, metrics=metrics, best=best, relres=relres)


def stats_accumulate(dst, src):
    """pg_stats_accumulate: dst += src in place (contiguous float64 CUDA tensors of one size); returns dst."""
    lib = L.load()
    if dst.numel() != src.numel():
        raise ValueError("dst and src must have the same number of entries")
    L.check(lib.pg_stats_accumulate(L.ptr(dst), L.ptr(src.contiguous()), dst.numel(), L.stream_ptr()))
    return dst


def basic_library_rows(u, u_x, u_y, lap_u):
    """pg_basic_library_rows: Theta [N][6] = [1, u, u_x, u_y, lap, u^2] (basic:75-101) from four arrays of N values."""
    torch = L.torch_cuda()
    lib = L.load()
    f = [_dev(np.asarray(a, dtype=np.float64).reshape(-1), torch.float64) for a in (u, u_x, u_y, lap_u)]
    n = f[0].numel()
    if any(x.numel() != n for x in f):
        raise ValueError("u, u_x, u_y, lap_u must have the same number of values")
    Theta = torch.empty((n, 6), dtype=torch.float64, device=f[0].device)
    L.check(lib.pg_basic_library_rows(L.ptr(f[0]), L.ptr(f[1]), L.ptr(f[2]), L.ptr(f[3]), n, L.ptr(Theta), L.stream_ptr()))
    return Theta


def fd_block_rows(U, d0, d1, dt, *, dialect, library, block):
    """pg_fd_block_rows: the block-mean rows [nrows][p + 1] (y first), nothing dropped."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, A0, A1 = U.shape
    p = L.LIB_WIDTH[library]
    bt, b0, b1 = (int(b) for b in block)
    Tr, R0, R1 = row_space(U.shape, dialect)
    nrows = -(-Tr // bt) * -(-R0 // b0) * -(-R1 // b1)
    rows = torch.empty((nrows, p + 1), dtype=torch.float64, device=U.device)
    L.check(lib.pg_fd_block_rows(L.ptr(U), T, A0, A1, float(d0), float(d1), float(dt), dialect, library, bt, b0, b1,
                                 L.ptr(rows), L.stream_ptr()))
    return rows


def rows_residual_ss(X, y, coef, *, fold_of_row=None, eval_fold=-1):
    """pg_rows_residual_ss: exact held-out residual sums of J fitted models from rows on the device.
    X [n][p], y [n], coef [J][p] -> (ss [J] NumPy, n_rows)."""
    torch = L.torch_cuda()
    lib = L.load()
    X = _dev(X, torch.float64)
    y = _dev(y, torch.float64)
    n, p = X.shape
    coef = _dev(np.asarray(coef, dtype=np.float64) if not isinstance(coef, torch.Tensor) else coef, torch.float64).reshape(-1, p).contiguous()
    fr = _dev(fold_of_row, torch.uint8)
    out = []
    for j0 in range(0, coef.shape[0], 32):
        c = coef[j0:j0 + 32].contiguous()
        ss = torch.empty(c.shape[0] + 1, dtype=torch.float64, device=X.device)
        L.check(lib.pg_rows_residual_ss(L.ptr(X), L.ptr(y), n, p, p, L.ptr(fr), int(eval_fold), L.ptr(c), c.shape[0], L.ptr(ss),
                                        L.stream_ptr()))
        out.append(ss)
    host = [o.cpu().numpy() for o in out]
    return np.concatenate([h[:-1] for h in host]), int(host[0][-1])


def fd_residual_ss(U, d0, d1, dt, coef, *, dialect, library, block=(1, 1, 1), fold_of_row=None, fold_of_frame=None,
                   n_folds=1, eval_fold=-1):
    """pg_fd_residual_ss: the same straight from the field (the block rows are re-formed on the fly; a second pass
    used only when the statistics-derived residual cancels).  coef [J][p] -> (ss [J] NumPy, n_rows)."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, A0, A1 = U.shape
    p = L.LIB_WIDTH[library]
    bt, b0, b1 = (int(b) for b in block)
    coef = _dev(np.asarray(coef, dtype=np.float64) if not isinstance(coef, torch.Tensor) else coef, torch.float64).reshape(-1, p).contiguous()
    fr = _dev(fold_of_row, torch.uint8)
    ff = _dev(fold_of_frame, torch.int32)
    out = []
    for j0 in range(0, coef.shape[0], 32):
        c = coef[j0:j0 + 32].contiguous()
        ss = torch.empty(c.shape[0] + 1, dtype=torch.float64, device=U.device)
        L.check(lib.pg_fd_residual_ss(L.ptr(U), T, A0, A1, float(d0), float(d1), float(dt), dialect, library, bt, b0, b1,
                                      L.ptr(fr), L.ptr(ff), n_folds, int(eval_fold), L.ptr(c), c.shape[0], L.ptr(ss),
                                      L.stream_ptr()))
        out.append(ss)
    host = [o.cpu().numpy() for o in out]
    return np.concatenate([h[:-1] for h in host]), int(host[0][-1])


def ks_rollout(U, d0, d1, dt, coef, n_steps, *, library):
    """pg_ks_rollout: explicit-Euler rollout of the fitted PDE from frame 0 (ks2d:1804-1838) -> rmse [n_steps]."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, A0, A1 = U.shape
    coef = _dev(np.asarray(coef, dtype=np.float64) if not isinstance(coef, torch.Tensor) else coef, torch.float64).reshape(-1)
    if coef.numel() != L.LIB_WIDTH[library]:
        raise ValueError("one coefficient per library column")
    n_steps = int(n_steps)
    work = torch.empty((2, A0, A1), dtype=torch.float64, device=U.device)
    rmse = torch.empty((max(n_steps, 0),), dtype=torch.float64, device=U.device)
    L.check(lib.pg_ks_rollout(L.ptr(U), T, A0, A1, float(d0), float(d1), float(dt), library, L.ptr(coef), n_steps,
                              L.ptr(work), L.ptr(rmse) if n_steps > 0 else None, L.stream_ptr()))
    return rmse


def ar_rollout_sums(U, d0, d1, dt, term_ids, coef, k_steps, t0, t1, mask=None):
    """pg_ar_rollout: (sum e^2, sum y, sum y^2, count) of the k-step Euler rollouts from every start frame of [t0, t1)."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, H, W = U.shape
    ids = _dev(np.asarray(term_ids, dtype=np.int32))
    cf = _dev(np.asarray(coef, dtype=np.float64).reshape(-1))
    if ids.numel() != cf.numel():
        raise ValueError("one coefficient per term")
    m = None if mask is None else _dev(np.ascontiguousarray(np.asarray(mask, dtype=bool)).astype(np.uint8))
    if m is not None and tuple(m.shape) != (H, W):
        raise ValueError(f"spatial_mask shape {tuple(m.shape)} does not match field shape {(H, W)}")
    n_start = int(t1) - int(k_steps) - int(t0)
    work = torch.empty((2, max(n_start, 1), H, W), dtype=torch.float64, device=U.device)
    out = torch.empty(4, dtype=torch.float64, device=U.device)
    L.check(lib.pg_ar_rollout(L.ptr(U), T, H, W, float(d0), float(d1), float(dt), L.ptr(ids), L.ptr(cf), ids.numel(),
                              int(k_steps), int(t0), int(t1), L.ptr(m), L.ptr(work), L.ptr(out), L.stream_ptr()))
    return out.cpu().numpy()


def one_step_sums(u_field, ut_pred, dt, mask=None):
    """pg_one_step_ss: (sum (u[t+1] - (u[t] + dt ut_pred[t]))^2, count) over t < min(len(u) - 1, len(ut_pred))."""
    torch = L.torch_cuda()
    lib = L.load()
    u = _dev(np.asarray(u_field, dtype=np.float64) if not isinstance(u_field, torch.Tensor) else u_field, torch.float64)
    ut = _dev(np.asarray(ut_pred, dtype=np.float64) if not isinstance(ut_pred, torch.Tensor) else ut_pred, torch.float64)
    t_max = min(u.shape[0] - 1, ut.shape[0])
    if t_max <= 0:
        return None
    frame = int(np.prod(u.shape[1:]))
    if int(np.prod(ut.shape[1:])) != frame:
        raise ValueError("u_field and ut_pred must have the same frame shape")
    m = None if mask is None else _dev(np.ascontiguousarray(np.asarray(mask, dtype=bool)).astype(np.uint8))
    out = torch.empty(2, dtype=torch.float64, device=u.device)
    L.check(lib.pg_one_step_ss(L.ptr(u), L.ptr(ut), t_max, frame, float(dt), L.ptr(m), L.ptr(out), L.stream_ptr()))
    return out.cpu().numpy()


def fit_metric_sums(y_true, y_pred):
    """pg_fit_metrics: the ten sums behind rmse / r2 / mae / std / corr of (y_true, y_pred); NumPy array [10]."""
    torch = L.torch_cuda()
    lib = L.load()
    yt = _dev(np.asarray(y_true, dtype=np.float64).ravel() if not isinstance(y_true, torch.Tensor) else y_true.reshape(-1), torch.float64)
    yp = _dev(np.asarray(y_pred, dtype=np.float64).ravel() if not isinstance(y_pred, torch.Tensor) else y_pred.reshape(-1), torch.float64)
    if yt.numel() != yp.numel():
        raise ValueError("y_true and y_pred must have the same number of elements")
    out = torch.zeros(10, dtype=torch.float64, device=yt.device)
    if yt.numel() == 0:
        return out.cpu().numpy(), 0
    L.check(lib.pg_fit_metrics(L.ptr(yt), L.ptr(yp), yt.numel(), L.ptr(out), L.stream_ptr()))
    return out.cpu().numpy(), yt.numel()


def rows_metrics_batched(X, y, coef, want_resid=False):
    """pg_rows_metrics_batched: X [B][n][p], y [B][n], coef [B][p] -> sums [B][10] (as fit_metric_sums) (+ resid [B][n])."""
    torch = L.torch_cuda()
    lib = L.load()
    X = _dev(X, torch.float64)
    y = _dev(y, torch.float64)
    B, n, p = X.shape
    coef = _dev(coef, torch.float64).reshape(B, p).contiguous()
    sums = torch.empty((B, 10), dtype=torch.float64, device=X.device)
    resid = torch.empty((B, n), dtype=torch.float64, device=X.device) if want_resid else None
    L.check(lib.pg_rows_metrics_batched(L.ptr(X), L.ptr(y), L.ptr(coef), B, n, p, p, L.ptr(sums), L.ptr(resid), L.stream_ptr()))
    return (sums, resid) if want_resid else sums


def gaussian_taps(sigma, truncate=4.0):
    """scipy.ndimage's normalised Gaussian taps (_gaussian_kernel1d, order 0): (radius, weights [2 r + 1])."""
    sigma = float(sigma)
    r = int(truncate * sigma + 0.5)
    x = np.arange(-r, r + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return r, np.ascontiguousarray(phi[::-1])


def gaussian_filter_frames(U, sigma, truncate=4.0, two_pass=False):
    """scipy.ndimage.gaussian_filter(frame, sigma) (mode="reflect") of every frame of a float32 / float64 stack: rows,
    then columns, rounding to the stack's dtype in between, as scipy does -- one fused pass through shared memory
    (pg_reflect_gauss2d) for radius <= 32, else (or with ``two_pass``) two pg_reflect_conv passes; same bits."""
    torch = L.torch_cuda()
    lib = L.load()
    if isinstance(U, np.ndarray):
        if U.dtype not in (np.float32, np.float64):
            U = U.astype(np.float64)
        U = _dev(U)
    if U.dtype not in (torch.float32, torch.float64):
        U = U.to(torch.float64)
    U = U.contiguous()
    T, A0, A1 = U.shape
    if float(sigma) <= 0:
        return U.clone()
    r, w = gaussian_taps(sigma, truncate)
    w_d = _dev(w)
    out = torch.empty_like(U)
    dt = 0 if U.dtype == torch.float32 else 1
    if r <= 32 and not two_pass:
        L.check(lib.pg_reflect_gauss2d(L.ptr(U), dt, T, A0, A1, L.ptr(w_d), r, L.ptr(out), L.stream_ptr()))
        return out
    tmp = torch.empty_like(U)
    L.check(lib.pg_reflect_conv(L.ptr(U), dt, T, A0, A1, 0, L.ptr(w_d), r, L.ptr(tmp), L.stream_ptr()))
    L.check(lib.pg_reflect_conv(L.ptr(tmp), dt, T, A0, A1, 1, L.ptr(w_d), r, L.ptr(out), L.stream_ptr()))
    return out


def time_moving_average(U, window):
    """pg_time_moving_average: reflect-padded moving average along t (ks2d:145-161)."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, A0, A1 = U.shape
    out = torch.empty_like(U)
    L.check(lib.pg_time_moving_average(L.ptr(U), T, A0, A1, int(window), L.ptr(out), L.stream_ptr()))
    return out


def periodic_gaussian_taps(n, sigma_px):
    """Taps (offsets, weights) of the periodic Gaussian g = ifft(exp(-sigma^2 k^2 / 2)) on n points
    (ks2d:135-138).  When the transfer function is below 1e-17 at the Nyquist frequency (sigma >= 2.82 px) g is
    numerically the wrapped Gaussian and is cut where it falls below 1e-17 of its peak; for smaller sigma the
    spectral cut-off makes g ring with slowly decaying tails and all n taps are kept."""
    sigma = float(sigma_px)
    k = 2.0 * np.pi * np.fft.fftfreq(n)
    g = np.fft.ifft(np.exp(-0.5 * sigma ** 2 * k ** 2)).real
    r = int(np.ceil(sigma * np.sqrt(2.0 * np.log(1e17)))) + 1
    if np.exp(-0.5 * sigma ** 2 * np.pi ** 2) < 1e-17 and 2 * r + 1 < n:
        off = np.arange(-r, r + 1)
    else:
        off = np.arange(n)
    return off.astype(np.int32), g[off % n].astype(np.float64)


def gaussian_smooth_periodic(U, sigma_px, method="auto"):
    """gaussian_smooth_periodic_2d (ks2d:125-142) for every frame of U.  ``method``: "conv" = two passes of
    pg_periodic_conv (a direct circular convolution with the taps of the periodic Gaussian: no FFT, but all n taps per
    axis when sigma < 2.82 px and at least 53 otherwise), "fft" = pg_periodic_gaussian_fft (the reference's own
    formulation, cuFFT), "auto" = the FFT for frames of 128 x 128 and more (256 x 2048^2, sigma = 3: 29 ms against
    226 ms; sigma = 1, where the direct route needs all 2048 taps: 29 ms against ~11 s), the direct route below."""
    torch = L.torch_cuda()
    lib = L.load()
    U = field(U)
    T, A0, A1 = U.shape
    if float(sigma_px) <= 0:
        return U.clone()
    if method not in ("auto", "conv", "fft"):
        raise ValueError("method must be 'auto', 'conv' or 'fft'")
    if method == "auto":
        method = "fft" if (A0 >= 128 and A1 >= 128) else "conv"
    if method == "fft":
        out = torch.empty_like(U)
        L.check(lib.pg_periodic_gaussian_fft(L.ptr(U), T, A0, A1, float(sigma_px), L.ptr(out), L.stream_ptr()))
        return out
    tmp, out = torch.empty_like(U), torch.empty_like(U)
    for axis, src, dst, n in ((0, U, tmp, A0), (1, tmp, out, A1)):
        off, w = periodic_gaussian_taps(n, sigma_px)
        off_d, w_d = _dev(off), _dev(w)
        L.check(lib.pg_periodic_conv(L.ptr(src), T, A0, A1, axis, L.ptr(off_d), L.ptr(w_d), len(off), L.ptr(dst),
                                     L.stream_ptr()))
    return out


def synth_field(T, A0, A1, *, t_offset=0, T_total=None, seed=0, kind=0, noise=0.0, out=None):
    """pg_synth_field: synthetic benchmark stack generated in HBM (no host transfer)."""
    torch = L.torch_cuda()
    lib = L.load()
    if out is None:
        out = torch.empty((T, A0, A1), dtype=torch.float64, device="cuda")
    L.check(lib.pg_synth_field(L.ptr(out), T, A0, A1, int(t_offset), int(T_total if T_total is not None else T),
                               int(seed), int(kind), float(noise), L.stream_ptr()))
    return out
