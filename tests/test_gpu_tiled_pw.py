"""GPU parity of the tiled pointwise kernel (K1 with every grid point a row, tiled_pw.cu): against the
generic reference-arithmetic kernel and against the oracle's rows.  Layouts exercise the periodic wrap on
every side, several tiles, a ragged last tile row (extents that are not multiples of 48), basic_usage's
interior masks on all four borders, frame chunking across persistent CTAs, time-holdout folds (up to 5),
out-of-range fold ids and the non-finite fallback."""

import numpy as np
import pytest

from helpers import assert_stats_close, ks_rows
from oracle import basic as OB
from oracle import gram

pytestmark = pytest.mark.gpu

KS_LIBS = {"LIB_KS_TRUE": ("true", False), "LIB_KS_TRUE_ADV": ("true", True), "LIB_KS_RICH": ("rich", False),
           "LIB_KS_RICH_NOADV": (None, None)}


@pytest.fixture(scope="module")
def env():
    from pde_b200 import _lib as L
    from pde_b200 import ops

    return L, ops


def field(ops, shape, seed, kind=0):
    return ops.synth_field(*shape, seed=seed, kind=kind, noise=0.05)


def basic_rows(U, d0, d1, dt):
    """Oracle rows of the basic_usage dialect; the C ABI's (d0, d1) are (dy, dx) (basic:58-59)."""
    d = OB.compute_derivatives(U, d1, d0, dt)
    return OB.build_library(*d[1:])[0], d[0].reshape(-1)


@pytest.mark.parametrize("libname", list(KS_LIBS))
@pytest.mark.parametrize("shape", [(5, 48, 128),      # one tile, wraps on all four sides
                                   (4, 96, 256),      # 2 x 2 tiles
                                   (6, 64, 128),      # ragged last tile row (64 = 48 + 16): wrap rows inside the box
                                   (3, 20, 128),      # a single ragged tile: both wraps in one tile
                                   (4, 100, 208),     # ragged rows; shifted last tile column (208 = 128 + 80), masked columns
                                   (3, 50, 330),      # width % 16 == 10: the shifted column goes through the second tensor map
                                   (3, 49, 128),      # the frame ends one row below a whole tile: its second halo row wraps
                                   (3, 97, 256)])
@pytest.mark.parametrize("d0,d1", [(0.5, 0.4), (0.5, 0.5)])      # square cells take the rho == 1 instantiation
def test_ks_pointwise_tiled(env, libname, shape, d0, d1):
    L, ops = env
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    U = field(ops, shape, seed=shape[1])
    dt = 1e-3
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(1, 1, 1))
    gen = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()[0]
    til = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_TILED, **kw).cpu().numpy()[0]
    assert_stats_close(til, gen, p)
    dictionary, adv = KS_LIBS[libname]
    if dictionary is not None:
        _, X, y = ks_rows(U.cpu().numpy(), d0, d1, dt, dictionary, adv, (1, 1, 1))
        assert_stats_close(til, gram.pack_stats(X, y), p)


@pytest.mark.parametrize("shape", [(4, 48, 128),      # one tile: all four borders masked
                                   (5, 60, 144),      # ragged rows and columns (second tile column has 16 columns)
                                   (3, 100, 304),     # 3 x 3 tiles, interior tile unmasked
                                   (4, 50, 144),      # last tile row holds only border points
                                   (6, 7, 128)])      # fewer rows than one band
def test_basic_pointwise_tiled(env, shape):
    L, ops = env
    U = field(ops, shape, seed=shape[2], kind=1)
    d0, d1, dt = 0.3, 0.25, 0.1
    kw = dict(dialect=L.FD_BASIC_TRIM, library=L.LIB_BASIC)
    gen = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()[0]
    til = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_TILED, **kw).cpu().numpy()[0]
    assert_stats_close(til, gen, 6)
    X, y = basic_rows(U.cpu().numpy(), d0, d1, dt)
    assert til[0] == X.shape[0]
    assert_stats_close(til, gram.pack_stats(X, y), 6)


def test_pointwise_two_frames_and_eight_folds(env):
    """The shortest stack (one row frame) and the maximum number of folds."""
    L, ops = env
    U = field(ops, (2, 96, 128), seed=21)
    for dialect, lib, p in ((L.FD_KS_PERIODIC, L.LIB_KS_TRUE, 3), (L.FD_BASIC_TRIM, L.LIB_BASIC, 6)):
        gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, dialect=dialect, library=lib, variant=L.VARIANT_GENERIC).cpu().numpy()
        til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, dialect=dialect, library=lib, variant=L.VARIANT_TILED).cpu().numpy()
        assert_stats_close(til[0], gen[0], p)
    T = 33
    V = field(ops, (T, 48, 128), seed=22)
    fof = (np.arange(T - 1) // 4).astype(np.int32)           # 8 folds of 4 frames
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE_ADV, fold_of_frame=fof, n_folds=8)
    gen = ops.fd_lib_gram(V, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(V, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(8):
        assert til[f][0] == 4 * 48 * 128
        assert_stats_close(til[f], gen[f], 5)


def test_pointwise_time_folds_and_chunks(env):
    """A long thin stack is cut into many frame chunks; 5 time-holdout folds (fold id per frame) cost a
    flush per fold change; results are run-to-run bit-identical."""
    L, ops = env
    T = 161
    U = field(ops, (T, 48, 256), seed=9)
    fof = (np.arange(T - 1) * 5 // (T - 1)).astype(np.int32)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(1, 1, 1), fold_of_frame=fof, n_folds=5)
    a = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    b = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    assert np.array_equal(a, b)
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    _, X, y = ks_rows(U.cpu().numpy(), 0.5, 0.5, 1e-3, "true", False, (1, 1, 1))
    fold_of_point = np.repeat(fof, 48 * 256)
    for f in range(5):
        assert_stats_close(a[f], gen[f], 3)
        assert_stats_close(a[f], gram.pack_stats(X[fold_of_point == f], y[fold_of_point == f]), 3)
    # interleaved folds: a flush at every frame
    fof2 = (np.arange(T - 1) % 3).astype(np.int32)
    kw["fold_of_frame"], kw["n_folds"] = fof2, 3
    a = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    for f in range(3):
        assert_stats_close(a[f], gen[f], 3)


def test_pointwise_basic_folds(env):
    L, ops = env
    T = 40
    U = field(ops, (T, 100, 272), seed=3, kind=1)
    fof = (np.arange(T - 1) >= 27).astype(np.int32)
    kw = dict(dialect=L.FD_BASIC_TRIM, library=L.LIB_BASIC, fold_of_frame=fof, n_folds=2)
    til = ops.fd_lib_gram(U, 0.3, 0.3, 0.1, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    X, y = basic_rows(U.cpu().numpy(), 0.3, 0.3, 0.1)
    fold_of_point = np.repeat(fof, 96 * 268)
    for f in range(2):
        assert_stats_close(til[f], gram.pack_stats(X[fold_of_point == f], y[fold_of_point == f]), 6)


def test_pointwise_excluded_frames_are_skipped(env):
    L, ops = env
    U = field(ops, (7, 48, 128), seed=5)
    fof = np.array([0, 1, -7, 0, -1, 1], dtype=np.int32)   # frames 2 and 4 are excluded on purpose (negative ids)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(1, 1, 1), fold_of_frame=fof, n_folds=2)
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    assert til[0][0] == 2 * 48 * 128 and til[1][0] == 2 * 48 * 128
    for f in range(2):
        assert_stats_close(til[f], gen[f], 3)


def test_pointwise_nonfinite_falls_back_to_exact_drop(env):
    """A NaN poisons the fast kernel's accumulators; the conditional generic launch then reproduces the
    reference's drop-non-finite-rows semantics (ks2d:1633-1636) and reports the count."""
    L, ops = env
    U = field(ops, (5, 48, 128), seed=6)
    U[1, 10, 100] = float("nan")
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(1, 1, 1), return_nonfinite=True)
    gen, bad_g = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw)
    til, bad_t = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw)
    assert int(bad_g[0].item()) == int(bad_t[0].item()) > 0
    assert np.isfinite(til.cpu().numpy()).all()
    assert np.array_equal(til.cpu().numpy(), gen.cpu().numpy())
    Uh = U.cpu().numpy()
    _, X, y = ks_rows(Uh, 0.5, 0.5, 1e-3, "true", False, (1, 1, 1))
    assert X.shape[0] == 4 * 48 * 128 - int(bad_t[0].item())    # the oracle (ks2d:394-395) dropped the same rows
    assert_stats_close(til.cpu().numpy()[0], gram.pack_stats(X, y), 3)


def test_pointwise_unsupported_layouts_use_generic(env):
    import pde_b200

    L, ops = env
    U = field(ops, (4, 48, 128), seed=7)
    fold = np.zeros(3 * 48 * 128, dtype=np.uint8)
    with pytest.raises(pde_b200.PdeGramError, match="no tiled kernel"):
        ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, fold_of_row=fold,
                        variant=L.VARIANT_TILED)
    odd = field(ops, (4, 48, 135), seed=7)      # TMA and the 16-byte side cells need 16-byte aligned rows: even widths
    with pytest.raises(pde_b200.PdeGramError, match="no tiled kernel"):
        ops.fd_lib_gram(odd, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, variant=L.VARIANT_TILED)
    narrow = field(ops, (4, 48, 64), seed=7)
    with pytest.raises(pde_b200.PdeGramError, match="no tiled kernel"):
        ops.fd_lib_gram(narrow, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, variant=L.VARIANT_TILED)
