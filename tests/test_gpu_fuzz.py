"""Seeded random shape sweep of pg_fd_lib_gram: whatever kernel(s) the library picks (tiled, tiled + generic remainders,
two-stage sub-block rows) against the generic reference-arithmetic kernel on the same field.  Small extents on purpose:
they put the periodic wraps, ragged tile rows, shifted tile columns and ragged blocks next to each other."""

import numpy as np
import pytest

from helpers import assert_stats_close

pytestmark = pytest.mark.gpu

KS_LIBS = ["LIB_KS_TRUE", "LIB_KS_TRUE_ADV", "LIB_KS_RICH", "LIB_KS_RICH_NOADV"]


@pytest.fixture(scope="module")
def env():
    from pde_b200 import _lib as L
    from pde_b200 import ops

    return L, ops


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    kind = ["block", "block_big", "point_ks", "point_basic"][seed % 4]
    T = int(rng.integers(2, 14))
    A0 = int(rng.integers(4, 210))
    A1 = int(rng.integers(64, 420))
    if rng.random() < 0.25:
        A0 = int(rng.choice([47, 48, 49, 50, 63, 64, 65, 66, 95, 97, 127, 129, 193]))   # around whole tiles (48 / 64 rows)
    if rng.random() < 0.25:
        A1 = int(rng.choice([126, 128, 130, 136, 144, 254, 256, 258, 264, 384]))       # around whole tile columns
    if rng.random() < 0.8:
        A1 += A1 % 2                                    # mostly even widths (the tiled kernels' domain)
    if kind == "block":
        A0 = max(A0, 8)
        block = (int(rng.integers(1, 6)), 8, 8)
        if rng.random() < 0.7:
            A0, A1 = -(-A0 // 8) * 8, -(-A1 // 8) * 8   # mostly whole blocks
    elif kind == "block_big":
        A0, A1 = -(-max(A0, 8) // 8) * 8, -(-A1 // 8) * 8
        block = (int(rng.integers(1, 5)), 8 * int(rng.integers(1, 5)), 8 * int(rng.integers(1, 7)))
    else:
        A0 = max(A0, 5)
        block = (1, 1, 1)
    lib = "LIB_BASIC" if kind == "point_basic" else KS_LIBS[int(rng.integers(0, 4))]
    folds = ["none", "time", "row"][int(rng.integers(0, 2 if kind.startswith("point") else 3))]
    return kind, (T, A0, A1), block, lib, folds, rng


@pytest.mark.parametrize("seed", range(120))
def test_auto_matches_generic(env, seed):
    L, ops = env
    kind, shape, block, libname, folds, rng = _case(seed)
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    basic = kind == "point_basic"
    U = ops.synth_field(*shape, seed=seed, kind=1 if basic else 0, noise=0.05)
    kw = dict(dialect=L.FD_BASIC_TRIM if basic else L.FD_KS_PERIODIC, library=lib, block=block)
    nf = 1
    T, A0, A1 = shape
    if folds == "time":
        nf = int(rng.integers(2, 5))
        nbt = -(-(T - 1) // block[0])
        kw.update(fold_of_frame=np.repeat(rng.integers(0, nf, size=nbt), block[0])[:T - 1].astype(np.int32), n_folds=nf)
    elif folds == "row":
        nf = int(rng.integers(2, 4))
        nrows = -(-(T - 1) // block[0]) * -(-A0 // block[1]) * -(-A1 // block[2])
        kw.update(fold_of_row=rng.integers(0, nf, size=nrows).astype(np.uint8), n_folds=nf)
    d0, d1 = float(rng.uniform(0.2, 0.6)), float(rng.uniform(0.2, 0.6))
    gen = ops.fd_lib_gram(U, d0, d1, 1e-2, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    aut = ops.fd_lib_gram(U, d0, d1, 1e-2, variant=L.VARIANT_AUTO, **kw).cpu().numpy()
    for f in range(nf):
        if gen[f][0] == 0:
            assert aut[f][0] == 0
            continue
        assert_stats_close(aut[f], gen[f], p)
