"""Timing split of the patch-ensemble leg (BASELINE configs[2]): K2 rows, per-patch statistics, K3."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from pde_b200 import _lib as L  # noqa: E402
from pde_b200 import ops  # noqa: E402
from pde_b200 import patch as PP  # noqa: E402


def timed(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, out


def main():
    Tp, Hp = 500, 1024
    U32 = ops.synth_field(Tp, Hp, Hp, seed=1, kind=1, noise=0.02).float()
    _, t_train, t_test = PP.time_split(Tp, 2, 0.7)
    coords = PP.patch_grid(Hp, Hp, 21, 10)
    tr_pts, te_pts = PP.sample_patch_points(np.random.default_rng(0), coords, Hp, Hp, 21, 3, t_train, t_test, 120)
    B = tr_pts.shape[0]
    W6 = ops._dev(PP.poly_stencil(2, 3, 3, 1.0, 0.1, 0.1))
    tr_d, te_d = ops._dev(tr_pts.reshape(-1, 3)), ops._dev(te_pts.reshape(-1, 3))
    a_d, t_d = ops._dev(np.array([0.01])), ops._dev(np.array([1e-5]))
    ms_k2, (X, y) = timed(lambda: ops.poly_rows(U32, tr_d, W6, 2, 3, library=L.LIB_PATCH_FULL))
    ms_k2t, (Xt, yt) = timed(lambda: ops.poly_rows(U32, te_d, W6, 2, 3, library=L.LIB_PATCH_FULL))
    X, y, Xt, yt = X.view(B, 120, 8), y.view(B, 120), Xt.view(B, 40, 8), yt.view(B, 40)
    shift = X[:, 0, :].contiguous()
    ms_g, (st, mm) = timed(lambda: ops.rows_gram(X, y, shift=shift, want_minmax=True))
    ms_gt, se = timed(lambda: ops.rows_gram(Xt, yt, shift=shift))
    ms_k3, _ = timed(lambda: ops.stridge_batched(st[:, 0], 8, dialect=L.STRIDGE_SKLEARN, alphas=a_d, thresholds=t_d, max_iter=25,
                                                 colminmax=mm[:, 0], shift=shift, eval_stats=se[:, 0]))
    print(dict(patches=B, k2_train_ms=round(ms_k2, 3), k2_test_ms=round(ms_k2t, 3), gram_train_ms=round(ms_g, 3),
               gram_test_ms=round(ms_gt, 3), k3_ms=round(ms_k3, 3),
               k2_points_per_s=round(B * 160 / ((ms_k2 + ms_k2t) * 1e-3))))


if __name__ == "__main__":
    main()
