"""GPU parity of the tiled TMA kernel (K1, (bt,8,8) blocks): against the generic
reference-arithmetic kernel on the device and against the oracle's rows, over tile layouts that
exercise periodic wrap on every side, multi-tile halos, remainders handed to the generic kernel,
ragged last t-blocks, both fold mechanisms and every KS library."""

import numpy as np
import pytest

from helpers import assert_stats_close, ks_rows
from oracle import gram

pytestmark = pytest.mark.gpu

LIBS = {"LIB_KS_TRUE": ("true", False, 3), "LIB_KS_TRUE_ADV": ("true", True, 5), "LIB_KS_RICH": ("rich", False, 9),
        "LIB_KS_RICH_NOADV": (None, None, 7)}


@pytest.fixture(scope="module")
def env():
    from pde_b200 import _lib as L
    from pde_b200 import ops

    return L, ops


def field(ops, shape, seed):
    # smooth waves + 5 % noise, so every term has signal and the fourth derivative is noise-dominated
    return ops.synth_field(*shape, seed=seed, noise=0.05)


@pytest.mark.parametrize("libname", list(LIBS))
@pytest.mark.parametrize("shape,bt", [((7, 64, 128), 3),      # one tile: wraps on all four sides, two t-blocks
                                      ((10, 128, 256), 3),    # 2 x 2 tiles, three t-blocks
                                      ((9, 72, 144), 3),      # ragged tile row, shifted last tile column (144 = 128 + 16), ragged t
                                      ((7, 64, 136), 3),      # width % 16 == 8: the shifted column goes through the second tensor map
                                      ((8, 76, 328), 2),      # rows that are not whole blocks -> generic box; 328 = 2 x 128 + 72
                                      ((7, 64, 132), 3),      # ragged last block column (4 columns) -> generic box
                                      ((7, 65, 128), 3),      # the frame ends one row below a whole tile: its second halo row wraps
                                      ((5, 129, 256), 2),
                                      ((6, 72, 270), 2),      # 270 = 33 blocks + 6 columns: shifted tile at column 136 (second map)
                                      ((6, 64, 384), 1),      # bt = 1, three tiles in a row
                                      ((12, 192, 128), 5),    # bt = 5, three tiles in a column, ragged t
                                      ((7, 88, 128), 3),      # ragged last tile row (88 = 64 + 24): wrap rows inside the box
                                      ((7, 40, 256), 3),      # a single ragged tile row: top and bottom wrap in one tile
                                      ((7, 1080 // 8 * 8 // 5, 128), 3)])   # 216 rows = 3 x 64 + 24
def test_tiled_vs_generic_and_oracle(env, libname, shape, bt):
    L, ops = env
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    U = field(ops, shape, seed=shape[1] + bt)
    d0, d1, dt = 0.5, 0.4, 1e-3
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(bt, 8, 8))
    gen = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()[0]
    til = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_TILED, **kw).cpu().numpy()[0]
    assert_stats_close(til, gen, p)
    dictionary, adv, _ = LIBS[libname]
    if dictionary is not None:
        names, X, y = ks_rows(U.cpu().numpy(), d0, d1, dt, dictionary, adv, (bt, 8, 8))
        assert_stats_close(til, gram.pack_stats(X, y), p)


def test_tiled_folds(env):
    L, ops = env
    shape, bt = (10, 128, 256), 3
    U = field(ops, shape, seed=1)
    n_rows = 3 * 16 * 32
    fold = np.random.default_rng(0).integers(0, 2, size=n_rows).astype(np.uint8)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_RICH, block=(bt, 8, 8), n_folds=2)
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_row=fold, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_row=fold, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(2):
        assert_stats_close(til[f], gen[f], 9)
    names, X, y = ks_rows(U.cpu().numpy(), 0.5, 0.5, 1e-3, "rich", False, (bt, 8, 8))
    for f in range(2):
        assert_stats_close(til[f], gram.pack_stats(X[fold == f], y[fold == f]), 9)
    fof = (np.arange(shape[0] - 1) >= 6).astype(np.int32)
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_frame=fof, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_frame=fof, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(2):
        assert_stats_close(til[f], gen[f], 9)
    assert til[0][0] == 2 * 16 * 32 and til[1][0] == 16 * 32


def test_tiled_row_folds_on_ragged_tile_rows(env):
    L, ops = env
    U = field(ops, (10, 88, 256), seed=14)
    fold = np.random.default_rng(3).integers(0, 2, size=3 * 11 * 32).astype(np.uint8)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8), fold_of_row=fold, n_folds=2)
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(2):
        assert_stats_close(til[f], gen[f], 3)


def test_tiled_many_time_folds(env):
    """Time-holdout folds (one id per frame) are accumulated one fold at a time: 5 folds, an out-of-range
    id, and folds that change inside a persistent CTA's chunk."""
    L, ops = env
    T, bt = 46, 3
    U = field(ops, (T, 64, 256), seed=11)
    fof = (np.arange(T - 1) // 9).astype(np.int32)          # 0..4, changes at t-block boundaries
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(bt, 8, 8), fold_of_frame=fof, n_folds=5)
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(5):
        assert til[f][0] == gen[f][0] == 3 * 8 * 32
        assert_stats_close(til[f], gen[f], 3)
    fof[6:9] = -1                                            # one t-block excluded on purpose (negative id): skipped
    kw["fold_of_frame"] = fof
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    assert til[0][0] == 2 * 8 * 32
    for f in range(5):
        assert_stats_close(til[f], gen[f], 3)
    fof[6:9] = 9                                             # an id >= n_folds is a caller error: loud (NaN + counter [1])
    kw["fold_of_frame"] = fof
    for variant in (L.VARIANT_GENERIC, L.VARIANT_TILED):
        st, ctr = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=variant, return_nonfinite=True, **kw)
        assert np.isnan(st.cpu().numpy()).all() and int(ctr[1].item()) == 8 * 32
    bad_pw = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(1, 1, 1),
                             fold_of_frame=fof, n_folds=5)
    assert np.isnan(bad_pw.cpu().numpy()).all()
    # the rich library keeps its statistics spread over the lanes (no private copy): same flush path
    kw.update(library=L.LIB_KS_RICH, fold_of_frame=(np.arange(T - 1) // 9).astype(np.int32))
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(5):
        assert_stats_close(til[f], gen[f], 9)


def test_tiled_row_folds_single_fold_and_two_frames(env):
    """Degenerate layouts: fold_of_row with a single fold (all zeros), and the shortest stack that has one
    t-block (bt + 1 frames)."""
    L, ops = env
    U = field(ops, (4, 64, 128), seed=12)
    fold = np.zeros(8 * 16, dtype=np.uint8)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE_ADV, block=(3, 8, 8), fold_of_row=fold, n_folds=1)
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    assert til.shape == (1, 28)
    assert_stats_close(til[0], gen[0], 5)
    U2 = field(ops, (2, 128, 128), seed=13)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_RICH_NOADV, block=(1, 8, 8))
    gen = ops.fd_lib_gram(U2, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    til = ops.fd_lib_gram(U2, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    assert_stats_close(til[0], gen[0], 7)


def test_tiled_many_chunks_and_determinism(env):
    """A longer stack is cut into frame chunks across persistent CTAs; results are run-to-run
    bit-identical (fixed work assignment, fixed-order reduction: no floating-point atomics)."""
    L, ops = env
    U = field(ops, (100, 64, 256), seed=4)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8))
    a = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    b = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    assert np.array_equal(a, b)
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    assert_stats_close(a[0], gen[0], 3)


def test_tiled_nonfinite_and_unsupported(env):
    import pde_b200

    L, ops = env
    U = field(ops, (7, 64, 128), seed=2)
    U[2, 10, 100] = float("nan")
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8), return_nonfinite=True)
    gen, bad_g = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw)
    til, bad_t = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw)
    assert int(bad_g[0].item()) == int(bad_t[0].item()) > 0
    assert_stats_close(til.cpu().numpy()[0], gen.cpu().numpy()[0], 3)
    with pytest.raises(pde_b200.PdeGramError, match="no tiled kernel"):
        ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 4, 8),
                        variant=L.VARIANT_TILED)
    for shape in [(4, 64, 131), (4, 64, 120)]:      # odd width (rows not 16-byte aligned); narrower than one tile
        with pytest.raises(pde_b200.PdeGramError, match="no tiled kernel"):
            ops.fd_lib_gram(field(ops, shape, seed=2), 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE,
                            block=(3, 8, 8), variant=L.VARIANT_TILED)


def test_shifted_tile_column_needs_no_generic_launch(env):
    """A width that is a multiple of 8 but not of 128 is covered by the tiled kernel alone (last tile column shifted
    left over its neighbour): the call launches exactly what a 128-multiple width launches."""
    L, ops = env
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8), variant=L.VARIANT_AUTO)
    counts = []
    for A1 in (256, 200, 1000):
        U = field(ops, (7, 64, A1), seed=3)
        n0 = L.load().pg_launch_count()
        ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, **kw)
        counts.append(L.load().pg_launch_count() - n0)
    assert counts[0] == counts[1] == counts[2]


@pytest.mark.parametrize("libname", ["LIB_KS_TRUE", "LIB_KS_RICH"])
@pytest.mark.parametrize("shape,block", [((10, 128, 256), (3, 16, 16)),     # 2 x 2 sub-blocks per block
                                         ((9, 72, 144), (2, 8, 32)),        # ragged edge blocks: 144 = 4 x 32 + 16
                                         ((7, 192, 384), (3, 64, 128)),     # a block as large as a tile
                                         ((8, 200, 272), (3, 24, 40))])     # multiples of 8 that do not divide the tile
def test_blocks_of_whole_sub_blocks_two_stage(env, libname, shape, block):
    """(bt, 8m, 8n) blocks: the tiled kernel writes the (bt, 8, 8) sub-block rows, the second stage averages them per
    block (ks2d:358-401 keeps ragged edge blocks with their own divisor) -- against the generic kernel and the oracle."""
    L, ops = env
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    U = field(ops, shape, seed=shape[2] + block[1])
    d0, d1, dt = 0.5, 0.4, 1e-3
    nrows = -(-(shape[0] - 1) // block[0]) * -(-shape[1] // block[1]) * -(-shape[2] // block[2])
    fold = np.random.default_rng(5).integers(0, 2, size=nrows).astype(np.uint8)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=block, fold_of_row=fold, n_folds=2)
    n0 = L.load().pg_launch_count()
    til = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    assert L.load().pg_launch_count() - n0 == 3          # tiled (rows), generic (combine), reduction
    gen = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    for f in range(2):
        assert_stats_close(til[f], gen[f], p)
    dictionary, adv, _ = LIBS[libname]
    _, X, y = ks_rows(U.cpu().numpy(), d0, d1, dt, dictionary, adv, block)
    assert X.shape[0] == nrows
    for f in range(2):
        assert_stats_close(til[f], gram.pack_stats(X[fold == f], y[fold == f]), p)
    # time folds and a NaN: the block that holds it is dropped and counted, like the generic kernel does
    U[2, 5, 7] = float("nan")
    fof = (np.arange(shape[0] - 1) >= (shape[0] - 1) // 2).astype(np.int32)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=block, fold_of_frame=fof, n_folds=2, return_nonfinite=True)
    til, bad_t = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_TILED, **kw)
    gen, bad_g = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_GENERIC, **kw)
    assert int(bad_t[0].item()) == int(bad_g[0].item()) > 0
    for f in range(2):
        assert_stats_close(til.cpu().numpy()[f], gen.cpu().numpy()[f], p)


def test_many_row_folds_go_through_the_sub_block_rows(env):
    """More than two per-row folds with (bt, 8, 8) blocks: rows from the tiled kernel, folds applied by the second stage."""
    L, ops = env
    shape, bt = (10, 128, 256), 3
    U = field(ops, shape, seed=11)
    fold = np.random.default_rng(2).integers(0, 5, size=3 * 16 * 32).astype(np.uint8)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(bt, 8, 8), fold_of_row=fold, n_folds=5)
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    for f in range(5):
        assert_stats_close(til[f], gen[f], 3)


@pytest.mark.parametrize("libname,shape,block", [("LIB_KS_TRUE", (10, 128, 256), (3, 8, 8)),
                                                 ("LIB_KS_RICH", (9, 72, 200), (2, 8, 8)),        # shifted tile column, ragged t
                                                 ("LIB_KS_TRUE_ADV", (7, 64, 136), (3, 8, 8)),
                                                 ("LIB_KS_TRUE", (8, 128, 256), (3, 16, 32))])    # two-stage blocks
def test_trailing_block_means_stand_in_for_the_trailing_frame(env, libname, shape, block):
    """pg_fd_lib_gram_tail: with the (8, 8) block means of the trailing frame the frame itself is a placeholder (time
    slabs whose trailing frame lives on another GPU send 1/64 of its bytes)."""
    import pde_b200

    L, ops = env
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    U = field(ops, shape, seed=21)
    fof = (np.arange(shape[0] - 1) >= (shape[0] - 1) // 2).astype(np.int32)
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=block, fold_of_frame=fof, n_folds=2)
    ref = ops.fd_lib_gram(U, 0.5, 0.4, 1e-3, **kw).cpu().numpy()
    means = ops.frame_block_means(U[-1])
    np.testing.assert_allclose(means.cpu().numpy(),
                               U[-1].cpu().numpy().reshape(shape[1] // 8, 8, shape[2] // 8, 8).mean(axis=(1, 3)), rtol=1e-13, atol=1e-15)
    V = U.clone()
    V[-1] = float("nan")                         # the placeholder must not be read for values
    got, bad = ops.fd_lib_gram(V, 0.5, 0.4, 1e-3, trailing_block_means=means, return_nonfinite=True, **kw)
    assert int(bad[0].item()) == 0
    for f in range(2):
        assert_stats_close(got.cpu().numpy()[f], ref[f], p, rtol=1e-12)
    with pytest.raises(pde_b200.PdeGramError, match="trailing_block_means"):
        ops.fd_lib_gram(V, 0.5, 0.4, 1e-3, trailing_block_means=means, variant=L.VARIANT_GENERIC, **kw)
    if block == (3, 8, 8) and shape[2] == 256:
        W = field(ops, (shape[0], shape[1], 132), seed=3)     # ragged block column: not covered by tiles alone
        with pytest.raises(pde_b200.PdeGramError, match="trailing_block_means"):
            ops.fd_lib_gram(W, 0.5, 0.4, 1e-3, trailing_block_means=torch_zeros(ops, (shape[1] // 8) * (132 // 8)), **kw)


def torch_zeros(ops, n):
    import torch

    return torch.zeros(n, dtype=torch.float64, device="cuda")


@pytest.mark.parametrize("block,libname,shape", [((3, 8, 8), "LIB_KS_TRUE", (10, 128, 256)),
                                                 ((3, 8, 8), "LIB_KS_RICH", (9, 72, 144)),      # ragged tile row, shifted column
                                                 ((3, 16, 16), "LIB_KS_TRUE", (7, 128, 256)),   # EMIT instantiation (two-stage route)
                                                 ((1, 1, 1), "LIB_KS_TRUE", (6, 96, 256)),      # pointwise kernel
                                                 ((1, 1, 1), "LIB_KS_TRUE_ADV", (6, 50, 136)),  # pointwise, ragged tile row, shifted column
                                                 ((1, 1, 1), "LIB_BASIC", (6, 100, 264))])
def test_producer_warp_and_decoupled_kernels_agree(env, monkeypatch, block, libname, shape):
    """Both structures of the tiled kernels stay under parity: with the producer warp (warp-specialised, setmaxnreg) and
    without it (every warp copies its own halo cells, the last arrival re-arms the stage).  The work assignment and the
    order of every floating-point operation are the same, so the statistics must agree BIT FOR BIT, with time folds, per-row
    folds and no folds; one of them is also checked against the generic reference-arithmetic kernel."""
    L, ops = env
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    U = field(ops, shape, seed=11)
    basic = libname == "LIB_BASIC"
    kw = dict(dialect=L.FD_BASIC_TRIM if basic else L.FD_KS_PERIODIC, library=lib, block=block)
    fof = (np.arange(shape[0] - 1) >= shape[0] // 2).astype(np.int32)
    cases = [dict(), dict(fold_of_frame=fof, n_folds=2)]
    if block == (3, 8, 8):
        nb = -(-(shape[0] - 1) // 3) * (shape[1] // 8) * (shape[2] // 8)
        cases.append(dict(fold_of_row=np.random.default_rng(3).integers(0, 2, size=nb).astype(np.uint8), n_folds=2))
    for extra in cases:
        out = {}
        for ws in ("0", "1"):
            monkeypatch.setenv("PG_TILED_WS", ws)
            monkeypatch.setenv("PG_PW_WS", ws)
            out[ws] = ops.fd_lib_gram(U, 0.5, 0.4, 1e-3, variant=L.VARIANT_TILED, **kw, **extra).cpu().numpy()
        assert np.array_equal(out["0"], out["1"]), (block, libname, sorted(extra))
        monkeypatch.delenv("PG_TILED_WS")
        monkeypatch.delenv("PG_PW_WS")
        gen = ops.fd_lib_gram(U, 0.5, 0.4, 1e-3, variant=L.VARIANT_GENERIC, **kw, **extra).cpu().numpy()
        for f in range(gen.shape[0]):
            assert_stats_close(out["1"][f], gen[f], p)


@pytest.mark.parametrize("libname", ["LIB_KS_TRUE", "LIB_KS_RICH"])
@pytest.mark.parametrize("shape,bt", [((10, 128, 256), 3), ((9, 72, 144), 3), ((8, 76, 328), 2), ((12, 192, 128), 5)])
def test_two_stack_tiled_vs_generic_and_oracle(env, libname, shape, bt):
    """pg_fd_lib_gram_two (ks2d:1448-1468, --denoise-space-on features): the library from U, the time derivative from a
    second stack Uy.  (bt, 8, 8) blocks run through the tiled kernel over U with y from (8, 8) block sums of Uy's frames
    k bt; against the generic kernel on the same two stacks and against the oracle's rows, with and without folds."""
    from oracle import ks2d as O

    L, ops = env
    lib = getattr(L, libname)
    p = L.LIB_WIDTH[lib]
    U = field(ops, shape, seed=5)
    Uy = field(ops, shape, seed=6) * 0.7 + 0.3 * U          # a different stack, correlated with U
    d0, d1, dt = 0.5, 0.4, 1e-3
    kw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(bt, 8, 8))
    fof = (np.arange(shape[0] - 1) >= shape[0] // 2).astype(np.int32)
    for extra in (dict(), dict(fold_of_frame=fof, n_folds=2)):
        gen = ops.fd_lib_gram(U, d0, d1, dt, Uy=Uy, variant=L.VARIANT_GENERIC, **kw, **extra).cpu().numpy()
        til = ops.fd_lib_gram(U, d0, d1, dt, Uy=Uy, variant=L.VARIANT_TILED, **kw, **extra).cpu().numpy()
        for f in range(gen.shape[0]):
            assert_stats_close(til[f], gen[f], p)
    Uh, Uyh = U.cpu().numpy(), Uy.cpu().numpy()
    names, terms = (O.build_dictionary_true if libname == "LIB_KS_TRUE" else O.build_dictionary)(Uh[:-1], d0, d1)
    X, y = O.build_blockwise_dataset((Uyh[1:] - Uyh[:-1]) / dt, terms, names, block_t=bt, block_x=8, block_y=8)
    assert_stats_close(til[0] + til[1], gram.pack_stats(X, y), p)
    # the same stack twice must reproduce the one-stack statistics
    one = ops.fd_lib_gram(U, d0, d1, dt, variant=L.VARIANT_TILED, **kw).cpu().numpy()[0]
    two = ops.fd_lib_gram(U, d0, d1, dt, Uy=U, variant=L.VARIANT_TILED, **kw).cpu().numpy()[0]
    assert_stats_close(two, one, p)
    # pointwise rows of two stacks have no tiled kernel: the default route is the generic kernel, asking for the tiled one fails loudly
    pw = dict(dialect=L.FD_KS_PERIODIC, library=lib, block=(1, 1, 1))
    a = ops.fd_lib_gram(U, d0, d1, dt, Uy=Uy, **pw).cpu().numpy()[0]
    b = ops.fd_lib_gram(U, d0, d1, dt, Uy=Uy, variant=L.VARIANT_GENERIC, **pw).cpu().numpy()[0]
    assert np.array_equal(a, b)
    with pytest.raises(L.PdeGramError):
        ops.fd_lib_gram(U, d0, d1, dt, Uy=Uy, variant=L.VARIANT_TILED, **pw)
