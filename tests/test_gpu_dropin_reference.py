"""The drop-in on the REAL reference modules.

``baseline/_ref`` holds the reference's scripts/ and examples/ verbatim (staged by ``__graft_entry__.build()`` in the
build container, git-ignored, shipped to the GPU box by gpurun).  Each test imports the unmodified script, rebinds its
hot-path functions with ``pde_b200.patch_reference(module)`` and runs the script's OWN ``main()`` (or, for the patch
script whose main() needs TIFF files, the body of its per-patch loop through the module's globals); what main() prints
must equal what the unmodified reference printed when the goldens were made (tests/golden/ks2d_configs.json,
basic.npz, patch.npz).
"""

import contextlib
import io
import re
import sys
from unittest import mock

import numpy as np
import pytest

from oracle import refload

pytestmark = pytest.mark.gpu

needs_ref = pytest.mark.skipif(not refload.available("ks2d"), reason="baseline/_ref is not staged (build() where /root/reference exists)")


def _run_main(mod, argv, name):
    buf = io.StringIO()
    with mock.patch.object(sys, "argv", [name] + argv), contextlib.redirect_stdout(buf):
        mod.main()
    return buf.getvalue()


def _parse_ks(text):
    res = {}
    m = re.search(r"hyperparams:\n(\{.*\})", text)
    if m:
        res["hyper"] = eval(m.group(1), {"__builtins__": {}}, {})
    res["coeffs"] = {n: float(v) for n, v in re.findall(r"^\s+(\S+)\s*: ([+-]\d+\.\d+)$", text, flags=re.M)}
    m = re.search(r"(?:Sampled|Blockwise) dataset: X=\((\d+), (\d+)\)", text)
    res["X_shape"] = [int(m.group(1)), int(m.group(2))]
    m = re.search(r"Rollout RMSE over (\d+) steps: first=([\d.eE+-]+), last=([\d.eE+-]+), mean=([\d.eE+-]+)", text)
    res["rollout"] = [float(m.group(k)) for k in (2, 3, 4)] if m else None
    return res


@needs_ref
@pytest.mark.parametrize("tag", ["c1", "c1_sweep", "c2", "c2_rich_sweep", "c2_denoise_features", "c2_denoise_all_pointwise"])
def test_ks2d_main_runs_on_the_gpu_functions(tag, golden_configs):
    import pde_b200

    gold = golden_configs[tag]
    ks = refload.load("ks2d", fresh=True)
    done = pde_b200.patch_reference(ks)
    assert {"gradients", "laplacian", "build_dictionary", "build_dictionary_true", "build_blockwise_dataset", "stridge",
            "ensemble_stridge", "rmse", "r2_score"} <= set(done)
    assert ks.stridge is pde_b200.ks2d.stridge and ks.build_blockwise_dataset is pde_b200.ks2d.build_blockwise_dataset
    got = _parse_ks(_run_main(ks, gold["argv"], "ks2d_stridge_benchmark.py"))
    assert got["X_shape"] == gold["X_shape"]
    assert got["coeffs"] == gold["coeffs_printed"]                 # six printed decimals, every term
    h, g = got["hyper"], gold["hyper"]
    assert (h["alpha"], h["threshold"], h["n_active"]) == (g["alpha"], g["threshold"], g["n_active"])
    if tag.startswith("c1"):
        assert h["r2_test"] == g["r2_test"] == 1.0
        np.testing.assert_allclose(h["rmse_test"], g["rmse_test"], rtol=1e-3)    # 2e-11: the rounding noise of an exact fit
    else:
        # (the denoise runs go through the periodic Gaussian, a circular convolution here and an FFT product there)
        tol = 1e-6 if "denoise" in tag else 1e-8
        np.testing.assert_allclose(h["r2_test"], g["r2_test"], rtol=tol)
        np.testing.assert_allclose(h["rmse_test"], g["rmse_test"], rtol=tol)


@needs_ref
def test_ks2d_main_ensemble_branch_survives_the_rebinding():
    """--regression ensemble calls ensemble_stridge(..., use_huber=True) (ks2d:1707-1715): the Huber inner solve is out of
    scope, so the rebound name delegates that call to the module's own function instead of raising."""
    import pde_b200

    ks = refload.load("ks2d", fresh=True)
    text_ref = _run_main(ks, ["--regression", "ensemble", "--n-seconds", "0.3", "--n-sample", "4000"], "ks2d")
    pde_b200.patch_reference(ks)
    text_gpu = _run_main(ks, ["--regression", "ensemble", "--n-seconds", "0.3", "--n-sample", "4000"], "ks2d")
    a, b = _parse_ks(text_ref), _parse_ks(text_gpu)
    assert a["X_shape"] == b["X_shape"] and a["coeffs"].keys() == b["coeffs"].keys()
    for k in a["coeffs"]:
        assert abs(a["coeffs"][k] - b["coeffs"][k]) <= 2e-6, (k, a["coeffs"][k], b["coeffs"][k])


@needs_ref
def test_basic_usage_main_runs_on_the_gpu_functions(golden_basic):
    import pde_b200

    ba = refload.load("basic", fresh=True)
    ba.plt.subplots.return_value = (mock.MagicMock(), mock.MagicMock())     # main() unpacks fig, axes (basic:215)
    text_ref = _run_main(ba, [], "basic_usage.py")
    done = pde_b200.patch_reference(ba)
    assert set(done) == {"compute_derivatives", "build_library", "stridge_regression"}
    text_gpu = _run_main(ba, [], "basic_usage.py")
    def pick(t):
        return (re.search(r"u_t = (.*)\n", t).group(1), re.search(r"R²: ([\d.\-]+)", t).group(1),
                re.search(r"Library shape: (\(.*\))", t).group(1))

    assert pick(text_gpu) == pick(text_ref)
    eq = pick(text_gpu)[0]
    coef = golden_basic["default_coef"]
    for c in coef[np.abs(coef) > 1e-6]:
        assert f"{c:.4f}" in eq


@needs_ref
def test_patch_loop_body_through_the_rebound_module(golden_patch):
    """patch main() reads TIFF files that are not in the tree; its per-patch loop body (patch:395-429) is driven through
    the module's globals after rebinding, on the golden stack with the golden sampled points."""
    import pde_b200

    pa = refload.load("patch", fresh=True)
    done = pde_b200.patch_reference(pa)
    assert {"stridge", "build_dataset", "local_poly_derivatives", "Library", "patch_grid", "regression_metrics"} <= set(done)
    g = golden_patch
    U = g["U"]
    lib8 = pa.Library(names=["1", "u", "u_x", "u_y", "lap(u)", "u^2", "u*u_x", "u*u_y"])
    assert pa.patch_grid(30, 34, 9, 4) == [tuple(r) for r in g["patch_grid_30_34_9_4"].tolist()]
    for k, pts in enumerate(g["loop_train_pts"][:6]):
        X, y = pa.build_dataset(U, [tuple(int(v) for v in p) for p in pts], rt=2, rs=3, deg=3, dt=1.0, dx=0.1, dy=0.1, lib=lib8)
        c = pa.stridge(X, y, alpha=0.01, threshold=1e-5)
        ref = g["loop_C"][k]
        assert np.array_equal(c != 0, ref != 0)
        np.testing.assert_allclose(c, ref, rtol=1e-8, atol=0)
        m = pa.regression_metrics(y, X @ c)
        assert set(m) >= {"r2", "rmse", "mae", "nrmse", "corr"}
