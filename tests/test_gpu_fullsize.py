"""Parity at BASELINE.json's full single-GPU size (configs[3]: 2048 x 2048 x 1024 float64, 34.4 GB): a full-width
sub-range of the stack against the NumPy oracle, and the whole stack through properties that do not need the oracle
at that size: the tiled TMA kernels against the generic
reference-arithmetic kernel on the same stack (statistics to 1e-10), additivity over time slabs (what the
multi-GPU path relies on), run-to-run bit identity, and the selected model of the 5 x 6 sweep."""

import numpy as np
import pytest

from helpers import assert_coef_close, assert_stats_close

pytestmark = pytest.mark.gpu

T, A = 1024, 2048


@pytest.fixture(scope="module")
def env():
    import torch

    from pde_b200 import _lib as L
    from pde_b200 import ops

    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs 40 GB of free device memory")
    U = ops.synth_field(T, A, A, seed=0, noise=0.05)
    fof = (np.arange(T - 1) >= int(0.7 * (T - 1)) // 3 * 3).astype(np.int32)
    yield L, ops, U, fof
    del U
    torch.cuda.empty_cache()


def test_c4_blockwise_tiled_vs_generic_and_slabs(env):
    L, ops, U, fof = env
    kw = dict(dialect=L.FD_KS_PERIODIC, library=L.LIB_KS_TRUE, block=(3, 8, 8), n_folds=2)
    til = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_frame=fof, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    again = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_frame=fof, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    assert np.array_equal(til, again)                       # fixed work assignment, fixed-order reduction
    gen = ops.fd_lib_gram(U, 0.5, 0.5, 1e-3, fold_of_frame=fof, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    assert til[0][0] + til[1][0] == 341 * 256 * 256
    for f in range(2):
        assert_stats_close(til[f], gen[f], 3)
    # eight time slabs of whole t-blocks (+ one halo frame each) add up to the whole stack
    from pde_b200.slabs import slab_bounds

    total = np.zeros_like(til)
    for lo, hi in slab_bounds(T - 1, 3, 8):
        total += ops.fd_lib_gram(U[lo:hi + 1], 0.5, 0.5, 1e-3, fold_of_frame=fof[lo:hi], variant=L.VARIANT_TILED, **kw).cpu().numpy()
    for f in range(2):
        assert_stats_close(total[f], til[f], 3, rtol=1e-12)
    # the model selected by the sweep is the same from either kernel's statistics
    from pde_b200 import ks2d as K

    a = K.fit_from_stats(til[0], til[1], K.TRUE_NAMES, grid_search=True)
    b = K.fit_from_stats(gen[0], gen[1], K.TRUE_NAMES, grid_search=True)
    assert (a["alpha"], a["threshold"]) == (b["alpha"], b["threshold"])
    assert_coef_close(a["coeffs"], b["coeffs"], what="c4 sweep")


@pytest.mark.parametrize("lib,p,dictionary", [("LIB_KS_TRUE", 3, "true"), ("LIB_KS_RICH", 9, "rich")])
def test_c4_tiled_sub_range_against_the_oracle(env, lib, p, dictionary):
    """The tiled kernel at the full C4 width against the NumPy ORACLE (not against another kernel of this library): a
    time range in the middle of the stack, every one of the 2048 columns, 512 rows; the rows wrap periodically, so
    the sub-stack is compared as a periodic field of its own, which is what both sides compute."""
    from helpers import ks_rows
    from oracle import gram

    L, ops, U, fof = env
    sub = U[400:425, 768:1280, :].contiguous()
    got = ops.fd_lib_gram(sub, 0.5, 0.5, 1e-3, dialect=L.FD_KS_PERIODIC, library=getattr(L, lib), block=(3, 8, 8),
                          variant=L.VARIANT_TILED).cpu().numpy()[0]
    names, X, y = ks_rows(sub.cpu().numpy(), 0.5, 0.5, 1e-3, dictionary, False, (3, 8, 8))
    ref = gram.pack_stats(X, y)
    assert got[0] == ref[0] == 8 * 64 * 256
    assert_stats_close(got, ref, p)
    # the north star's wording taken literally: entrywise relative error of every entry that is not a cancelling sum
    from bench import stats_rel_err

    cs, ew = stats_rel_err(got, ref, p)
    assert cs <= 1e-10 and ew <= 1e-9, (cs, ew)


@pytest.mark.parametrize("dialect,lib,p", [("FD_KS_PERIODIC", "LIB_KS_TRUE", 3), ("FD_BASIC_TRIM", "LIB_BASIC", 6)])
def test_c4_pointwise_tiled_vs_generic(env, dialect, lib, p):
    L, ops, U, fof = env
    frames = 192                                            # the generic pointwise kernel runs at ~0.05 TB/s
    kw = dict(dialect=getattr(L, dialect), library=getattr(L, lib), fold_of_frame=fof[:frames - 1] * 0 + (np.arange(frames - 1) >= 130),
              n_folds=2)
    til = ops.fd_lib_gram(U[:frames], 0.5, 0.5, 1e-3, variant=L.VARIANT_TILED, **kw).cpu().numpy()
    gen = ops.fd_lib_gram(U[:frames], 0.5, 0.5, 1e-3, variant=L.VARIANT_GENERIC, **kw).cpu().numpy()
    side = A if p == 3 else A - 4
    assert til[0][0] == 130 * side * side and til[1][0] == (frames - 1 - 130) * side * side
    for f in range(2):
        assert_stats_close(til[f], gen[f], p)
