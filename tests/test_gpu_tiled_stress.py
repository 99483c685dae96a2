"""Mid-size shape sweep of the tiled kernels against the generic kernel: many tiles per persistent CTA, several
frame chunks, ragged tile rows / t-blocks, every library, both fold mechanisms (pacing, stage re-arming and the
per-warp side cells are exercised across item boundaries)."""

import numpy as np
import pytest

from helpers import assert_stats_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    from pde_b200 import _lib as L
    from pde_b200 import ops

    return L, ops


BLOCK_CASES = [  # shape, bt, library, folds ("none" | "time3" | "row2")
    ((61, 320, 384), 3, "LIB_KS_TRUE", "time3"),
    ((40, 1024, 1024), 3, "LIB_KS_TRUE", "row2"),
    ((35, 1080, 1920), 4, "LIB_KS_TRUE_ADV", "none"),
    ((26, 512, 640), 5, "LIB_KS_RICH", "time3"),
    ((17, 200, 256), 2, "LIB_KS_RICH_NOADV", "row2"),
    ((300, 64, 128), 1, "LIB_KS_TRUE", "time3"),
    ((31, 1000, 1000), 3, "LIB_KS_TRUE", "row2"),        # shifted last tile column, width % 16 == 8
    ((22, 480, 720), 3, "LIB_KS_RICH", "time3"),         # 720 = 5 x 128 + 80
]


@pytest.mark.parametrize("shape,bt,libname,folds", BLOCK_CASES)
def test_blockwise_sweep(env, shape, bt, libname, folds):
    L, ops = env
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    U = ops.synth_field(*shape, seed=shape[0], noise=0.05)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(bt, 8, 8))
    nf = 1
    if folds == "time3":
        kw.update(fold_of_frame=(np.arange(shape[0] - 1) * 3 // (shape[0] - 1)).astype(np.int32), n_folds=3)
        nf = 3
    elif folds == "row2":
        nrows = -(-(shape[0] - 1) // bt) * (shape[1] // 8) * (shape[2] // 8)
        kw.update(fold_of_row=np.random.default_rng(1).integers(0, 2, size=nrows).astype(np.uint8), n_folds=2)
        nf = 2
    gen = ops.fd_lib_gram(U, 0.5, 0.4, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.5, 0.4, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(nf):
        assert_stats_close(til[f], gen[f], p)


POINT_CASES = [  # shape, dialect, library, number of time folds
    ((40, 1024, 1024), "FD_KS_PERIODIC", "LIB_KS_TRUE", 2),
    ((25, 1080, 1920), "FD_BASIC_TRIM", "LIB_BASIC", 3),
    ((30, 500, 768), "FD_KS_PERIODIC", "LIB_KS_RICH_NOADV", 1),
    ((30, 333, 640), "FD_BASIC_TRIM", "LIB_BASIC", 1),
    ((150, 96, 256), "FD_KS_PERIODIC", "LIB_KS_TRUE_ADV", 4),
    ((21, 1000, 1000), "FD_KS_PERIODIC", "LIB_KS_TRUE", 2),     # shifted last tile column, width % 16 == 8
    ((18, 250, 602), "FD_BASIC_TRIM", "LIB_BASIC", 2),          # basic_usage: shifted AND border-masked last column
    ((16, 144, 170), "FD_KS_PERIODIC", "LIB_KS_RICH", 1),
]


@pytest.mark.parametrize("shape,dialect,libname,nf", POINT_CASES)
def test_pointwise_sweep(env, shape, dialect, libname, nf):
    L, ops = env
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    U = ops.synth_field(*shape, seed=shape[1], kind=1 if "BASIC" in dialect else 0, noise=0.05)
    kw = dict(dialect=getattr(L, dialect), library=lib, n_folds=nf)
    if nf > 1:
        kw["fold_of_frame"] = (np.arange(shape[0] - 1) * nf // (shape[0] - 1)).astype(np.int32)
    gen = ops.fd_lib_gram(U, 0.3, 0.25, 1e-2, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.3, 0.25, 1e-2, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(nf):
        assert til[f][0] == gen[f][0]
        assert_stats_close(til[f], gen[f], p)
