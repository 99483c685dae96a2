"""Rebind the hot-path functions of a loaded reference script to the GPU implementations.

The reference defines its hot-path functions in the same files as ``main()`` and resolves
them through module globals at call time (ks2d:1518,1542,1600,1718; basic:188-200;
patch:420-423), so replacing module attributes is a complete drop-in: no source edit.

    import ks2d_stridge_benchmark as m
    pde_b200.patch_reference(m)      # m.stridge, m.build_blockwise_dataset, ... now run on the GPU
    m.main()
"""

from __future__ import annotations

_KS = ("rmse", "r2_score", "standardize_transform", "gradients", "laplacian", "build_dictionary", "build_dictionary_true", "build_blockwise_dataset",
       "standardize_fit", "ridge_fit", "stridge", "stridge_sign_constrained", "ensemble_stridge",
       "gaussian_smooth_periodic_2d", "time_smooth_moving_average")
_BASIC = ("compute_derivatives", "build_library", "stridge_regression")
_PATCH = ("regression_metrics", "stridge", "local_poly_derivatives", "build_dataset", "Library", "patch_grid",
          "safe_sample_points")
# gaussian_filter is scipy's (imported by name into patch / analyze_results): pde_b200.patch.gaussian_filter takes the
# whole stack at once, the scripts call it per frame, so it is offered but not rebound.


def patch_reference(module, dialect: str | None = None):
    """Replace the hot-path callables of ``module`` in place; returns the list of names rebound.

    ``dialect`` is "ks2d", "basic" or "patch"; by default it is inferred from the functions the
    module defines."""
    from . import basic_usage, ks2d, patch

    if dialect is None:
        if hasattr(module, "build_blockwise_dataset"):
            dialect = "ks2d"
        elif hasattr(module, "stridge_regression"):
            dialect = "basic"
        elif hasattr(module, "local_poly_derivatives"):
            dialect = "patch"
        else:
            raise ValueError("cannot infer which reference script this module is")
    src, names = {"ks2d": (ks2d, _KS), "basic": (basic_usage, _BASIC), "patch": (patch, _PATCH)}[dialect]
    done = []
    for n in names:
        if hasattr(module, n):
            new = getattr(src, n)
            if n == "ensemble_stridge":
                new = _ensemble_wrapper(getattr(module, n), new)
            setattr(module, n, new)
            done.append(n)
    return done


def _ensemble_wrapper(original, gpu):
    """The script's main() always calls ensemble_stridge(..., use_huber=True, huber_delta=...) for
    ``--regression ensemble`` (ks2d:1707-1715).  The Huber inner solve iterates on raw residuals (IRLS with a
    median scale), which is not Gram-reducible and is out of scope: that call keeps running the
    module's own function, so the rebound script never crashes; the ridge ensemble (use_huber=False) runs on the GPU."""
    def ensemble_stridge(X, y, *args, use_huber=False, **kw):
        if use_huber:
            return original(X, y, *args, use_huber=True, **kw)
        return gpu(X, y, *args, use_huber=False, **kw)

    ensemble_stridge.__wrapped__ = original
    return ensemble_stridge
