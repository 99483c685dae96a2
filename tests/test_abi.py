"""CPU-side checks of the boundary: the C-ABI library loads, exports every symbol the header
declares, mirrors its constants in Python, and refuses to compute without a CUDA device."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

import pde_b200
from pde_b200 import _lib as L

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "pdegram.h").read_text()


@pytest.fixture(scope="module")
def lib():
    if not L.LIB_PATH.exists():
        pde_b200.build()
    return pde_b200.load()


def test_library_exports_every_declared_symbol(lib):
    declared = re.findall(r"^PG_API\s+[\w \*]+?\b(pg_\w+)\s*\(", HEADER, flags=re.M)
    assert len(declared) >= 12 and sorted(declared) == sorted(L.EXPORTS)
    raw = ctypes.CDLL(str(L.LIB_PATH))
    for name in declared:
        assert hasattr(raw, name), name


def test_python_constants_mirror_header(lib):
    defs = {k: int(v) for k, v in re.findall(r"^#define (PG_\w+) \(?(-?\d+)\)?\s", HEADER, flags=re.M)}
    assert lib.pg_version() == defs["PG_VERSION"]
    assert (L.PG_MAX_P, L.PG_MAX_FOLDS) == (defs["PG_MAX_P"], defs["PG_MAX_FOLDS"])
    for py, h in [("FD_KS_PERIODIC", "PG_FD_KS_PERIODIC"), ("FD_BASIC_TRIM", "PG_FD_BASIC_TRIM"),
                  ("LIB_KS_TRUE", "PG_LIB_KS_TRUE"), ("LIB_KS_TRUE_ADV", "PG_LIB_KS_TRUE_ADV"),
                  ("LIB_KS_RICH", "PG_LIB_KS_RICH"), ("LIB_KS_RICH_NOADV", "PG_LIB_KS_RICH_NOADV"),
                  ("LIB_BASIC", "PG_LIB_BASIC"), ("LIB_KS_GRAD", "PG_LIB_KS_GRAD"), ("LIB_KS_LAP", "PG_LIB_KS_LAP"),
                  ("LIB_PATCH_MODEL4", "PG_LIB_PATCH_MODEL4"), ("LIB_PATCH_FULL", "PG_LIB_PATCH_FULL"),
                  ("LIB_PATCH_DERIVS", "PG_LIB_PATCH_DERIVS"), ("STRIDGE_KS", "PG_STRIDGE_KS"),
                  ("STRIDGE_SKLEARN", "PG_STRIDGE_SKLEARN"), ("STRIDGE_BASIC", "PG_STRIDGE_BASIC"),
                  ("STRIDGE_RMS_PRESCALE", "PG_STRIDGE_RMS_PRESCALE"), ("VARIANT_AUTO", "PG_VARIANT_AUTO"),
                  ("VARIANT_GENERIC", "PG_VARIANT_GENERIC"), ("VARIANT_TILED", "PG_VARIANT_TILED")]:
        assert getattr(L, py) == defs[h], py
    for lib_id, w in L.LIB_WIDTH.items():
        assert lib.pg_library_width(lib_id) == w
    assert lib.pg_library_width(99) < 0 and b"unknown library" in lib.pg_last_error()
    from oracle import gram

    for p in (1, 3, 9, 16):
        assert L.stats_len(p) == gram.stats_len(p)


def test_argument_validation_needs_no_gpu(lib):
    """PG_EINVAL paths return before any CUDA call, so they can be exercised on the CPU box."""
    rc = lib.pg_fd_lib_gram(None, 4, 8, 8, 1.0, 1.0, 1.0, 0, 0, 1, 1, 1, None, None, 1, None, None, 0, None)
    assert rc == -1 and b"U is null" in lib.pg_last_error()
    rc = lib.pg_stridge_batched(None, 1, 99, 0, 0, None, 1, None, 1, 25, None, None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"p must be" in lib.pg_last_error()
    rc = lib.pg_block_means(1, 1, 2, 2, 2, 0, 1, 1, 1, None)  # dummy non-null pointers, block_t = 0
    assert rc == -1 and b"block sizes" in lib.pg_last_error()


def test_no_cpu_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present; this test pins the CPU-box behaviour")
    with pytest.raises(pde_b200.PdeGramError, match="no CUDA device"):
        pde_b200.ks2d.stridge(np.ones((4, 2)), np.ones(4))
    with pytest.raises(pde_b200.PdeGramError, match="no CUDA device"):
        pde_b200.basic_usage.compute_derivatives(np.ones((3, 6, 6)), 1.0, 1.0, 1.0)
    with pytest.raises(pde_b200.PdeGramError, match="no CUDA device"):
        pde_b200.patch.build_dataset(np.ones((6, 8, 8), np.float32), [(2, 3, 3)], 2, 3, 3, 1.0, 0.1, 0.1,
                                     pde_b200.patch.Library(names=pde_b200.patch.FULL_NAMES))


def test_product_never_imports_oracle():
    pkg = ROOT / "pde-discovery-laser-matter_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.h")):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", f.read_text(), flags=re.M), f


def test_host_logic_patch_grid_and_sampling(golden_patch):
    """Host-side logic of the patch dialect (RNG draw order is part of the contract)."""
    from pde_b200 import patch as P

    g = golden_patch
    assert np.array_equal(np.array(P.patch_grid(30, 34, 9, 4)), g["patch_grid_30_34_9_4"])
    assert len(P.patch_grid(1024, 1024, 21, 10)) == 8464
    _, t_train, t_test = P.time_split(14, 2, 0.7)
    tr, te = P.sample_patch_points(np.random.default_rng(0), P.patch_grid(30, 34, 11, 5), 30, 34, 11, 3, t_train, t_test, 40)
    assert np.array_equal(tr, g["loop_train_pts"]) and np.array_equal(te, g["loop_test_pts"])
    from oracle import patch as OP

    np.testing.assert_allclose(P.poly_stencil(2, 3, 3, 1.0, 0.1, 0.1), OP.poly_stencil(2, 3, 3, 1.0, 0.1, 0.1), rtol=1e-13, atol=1e-18)
    agg = P.stability_aggregate(g["loop_C"], 1e-5)
    for k in ("freq", "median", "q25", "q75", "sign_stability", "agg"):
        np.testing.assert_array_equal(agg[k], g[f"loop_{k}"])


def test_host_logic_ks_fold_split():
    from oracle import ks2d as O
    from pde_b200 import ks2d as K

    fold, perm = K.split_folds(1000, np.random.default_rng(0))
    rng = np.random.default_rng(0)
    tr, te, _ = O.split_and_scale(["a"], np.ones((1000, 1)), np.ones(1000), rng)
    assert np.array_equal(np.flatnonzero(fold == 0), np.sort(tr)) and np.array_equal(np.flatnonzero(fold == 1), np.sort(te))
    assert K.RICH_NAMES == O.RICH_NAMES and K.TRUE_NAMES == O.TRUE_NAMES
    assert (K.GRID_ALPHAS, K.GRID_THRESHOLDS) == (O.GRID_ALPHAS, O.GRID_THRESHOLDS)


def test_host_binding_reports_instead_of_raising():
    """_xfer.bind_host_to_gpu is a tuning aid of the host-streaming path: where it cannot act (no GPU, no NVML, a cpuset
    without the GPU's local CPUs) it says why and leaves the process's affinity alone."""
    import os

    from pde_b200 import _xfer

    before = os.sched_getaffinity(0)
    info = _xfer.bind_host_to_gpu()
    assert set(info) == {"bound", "cpus", "numa_node", "why"}
    import torch

    if not torch.cuda.is_available():
        assert info["bound"] is False and info["why"] and os.sched_getaffinity(0) == before
