"""The noise floor of the patch dialect's two reformulations, measured against the UNMODIFIED reference
(baseline/_ref, or /root/reference through oracle.refload.stage()):

  (a) local_poly_derivatives (patch:193-246) solves a 245 x 20 least-squares problem per point with np.linalg.lstsq;
      the design matrix is constant, so K2 applies the fixed stencil pinv(A)[rows] instead.  How far apart are the rows?
  (b) stridge (patch:78-98) runs scikit-learn's StandardScaler + Ridge on the rows; K3 solves the same standardised
      system from sufficient statistics.  How far apart are the coefficients, rows (a) included?

CPU only (NumPy oracle vs reference): python tools/patch_floor.py [--patches 40]
Measured in the build container (numpy 2.3.5, scikit-learn 1.9.0, 12 patches of a 30 x 64 x 64 float32 stack):
rows 9.0e-15 relative, coefficients 4.5e-12 relative, support identical -- four orders of magnitude below the 1e-8 the
parity tests ask of the GPU path (tests/test_gpu_basic_patch.py, tests/test_gpu_dropin_reference.py).
"""

import argparse
import json
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))

import numpy as np  # noqa: E402


def main():
    from helpers import synthetic_stack
    from oracle import patch as OP
    from oracle import refload

    ap = argparse.ArgumentParser()
    ap.add_argument("--patches", type=int, default=12)
    args = ap.parse_args()
    refload.stage()
    pa = refload.load("patch")
    U = synthetic_stack((30, 64, 64), seed=3)
    rt, rs, deg, dt, dx, dy = 2, 3, 3, 1.0, 0.1, 0.1
    lib_ref, lib_o = pa.Library(names=list(OP.FULL_NAMES)), OP.Library(names=list(OP.FULL_NAMES))
    coords = OP.patch_grid(64, 64, 21, 10)
    _, t_train, t_test = OP.time_split(30, rt, 0.7)
    samples = OP.sample_patch_points(np.random.default_rng(0), coords, 64, 64, 21, rs, t_train, t_test, 120)
    W = OP.poly_stencil(rt, rs, deg, dt, dx, dy)
    rows, coefs, same = 0.0, 0.0, True
    for tr, _ in samples[: args.patches]:
        pts = [tuple(int(v) for v in p) for p in tr]
        Xr, yr = pa.build_dataset(U, pts, rt=rt, rs=rs, deg=deg, dt=dt, dx=dx, dy=dy, lib=lib_ref)
        Xs, ys = OP.build_dataset_stencil(U, tr, rt, rs, deg, dt, dx, dy, lib_o, W)
        rows = max(rows, float((np.abs(Xr - Xs) / np.abs(Xr).max(axis=0)).max()), float(np.abs(yr - ys).max() / np.abs(yr).max()))
        cr = pa.stridge(Xr, yr, alpha=0.01, threshold=1e-5)
        cs = OP.stridge(Xs, ys, alpha=0.01, threshold=1e-5)
        same &= bool(np.array_equal(cr != 0, cs != 0))
        nz = cr != 0
        if nz.any():
            coefs = max(coefs, float((np.abs(cr - cs)[nz] / np.abs(cr[nz])).max()))
    print(json.dumps({"patches": min(args.patches, len(samples)), "rows_max_rel_lstsq_vs_stencil": rows,
                      "coef_max_rel_reference_vs_stencil_plus_statistics": coefs, "same_support": same}))


if __name__ == "__main__":
    main()
