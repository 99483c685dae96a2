import sys, os; sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np, torch
from pde_b200 import _lib as L, ops
from quick_bench import time_call
T, A = 256, 2048
U = ops.synth_field(T, A, A, seed=0, noise=0.05)
nrows = ((T - 1 + 2) // 3) * (A // 8) * (A // 8)
fold = torch.from_numpy((np.random.default_rng(0).random(nrows) >= 0.7).astype(np.uint8)).cuda()
fof = (torch.arange(T - 1) >= int(0.7 * (T - 1))).to(torch.int32).cuda()
for lib, nm in ((L.LIB_KS_TRUE, "true"), (L.LIB_KS_RICH, "rich")):
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(3, 8, 8), n_folds=2)
    for mode in ("time folds", "row folds fused", "row folds two-stage"):
        os.environ.pop("PG_ROWFOLD_TWO_STAGE", None)
        if mode == "row folds two-stage":
            os.environ["PG_ROWFOLD_TWO_STAGE"] = "1"
        k = dict(fold_of_frame=fof) if mode == "time folds" else dict(fold_of_row=fold)
        best, avg = time_call(lambda: ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, **kw, **k), iters=6)
        print(nm, mode, round(best, 3), "ms", round(8 * U.numel() / best / 1e6), "GB/s", flush=True)
a = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_row=fold, **kw).cpu().numpy()
os.environ.pop("PG_ROWFOLD_TWO_STAGE")
b = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_row=fold, **kw).cpu().numpy()
print("max rel diff two-stage vs fused", float(np.max(np.abs(a - b) / (np.abs(b) + 1e-300))))
