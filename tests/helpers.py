"""Shared assertions for the parity tests."""

import numpy as np

from oracle import gram

GRAM_RTOL = 1e-10   # north star: Gram entries within 1e-10 relative
COEF_RTOL = 1e-8    # north star: coefficients within 1e-8 relative


def stats_scales(s, p):
    """Natural magnitude of every statistics entry: the Cauchy-Schwarz bound sqrt(<a,a><b,b>) of
    the inner product it is.  "Relative" error of an entry is measured against this, because a
    cross term such as sum(u_x*u_y) can cancel to ~0 and has no meaningful own-magnitude."""
    n, sy, syy, sx, b, G = gram.unpack_stats(s, p)
    d = np.sqrt(np.maximum(np.diag(G), 0))
    iu = np.triu_indices(p)
    return np.concatenate([[n, np.sqrt(n * syy), syy], np.sqrt(n) * d, np.sqrt(syy) * d, np.outer(d, d)[iu]])


def assert_stats_close(got, ref, p, rtol=GRAM_RTOL):
    got, ref = np.asarray(got, dtype=np.float64).ravel(), np.asarray(ref, dtype=np.float64).ravel()
    assert got.shape == ref.shape == (gram.stats_len(p),)
    assert got[0] == ref[0], f"row count differs: {got[0]} vs {ref[0]}"
    err = np.abs(got - ref) / np.maximum(stats_scales(ref, p), 1e-300)
    assert err.max() <= rtol, f"max scaled stats error {err.max():.3e} at entry {err.argmax()}"


def assert_coef_close(got, ref, rtol=COEF_RTOL, what=""):
    got, ref = np.asarray(got), np.asarray(ref)
    assert np.array_equal(got != 0, ref != 0), f"support differs {what}: {got} vs {ref}"
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=0, err_msg=what)


def ks_rows(U, dx, dy, DT, dictionary="rich", include_advection=False, block=(1, 1, 1)):
    """Oracle rows (X, y) for a KS-dialect field and block size."""
    from oracle import ks2d

    Ut = (U[1:] - U[:-1]) / DT
    if dictionary == "true":
        names, terms = ks2d.build_dictionary_true(U[:-1], dx, dy, include_advection=include_advection)
    else:
        names, terms = ks2d.build_dictionary(U[:-1], dx, dy)
    X, y = ks2d.build_blockwise_dataset(Ut, terms, names, block_t=block[0], block_x=block[1], block_y=block[2])
    return names, X, y


def synthetic_stack(shape, seed=0):
    """Smooth laser-image-shaped float32 stack in [0, 1] (SURVEY 8d, config C3 input)."""
    from scipy.ndimage import gaussian_filter

    a = gaussian_filter(np.random.default_rng(seed).standard_normal(shape), sigma=(1, 2, 2))
    a = (a - a.min()) / (a.max() - a.min())
    return a.astype(np.float32)


def rollout_case(tag, golden_configs, ks_default_stack):
    """Inputs of the rollout check of main() for one reference config: the (noisy) stack the script observed
    and the full-precision coefficients of the configuration it selected."""
    from oracle import ks2d

    U, dx, dy, DT = ks_default_stack
    Uo = U if tag == "c1" else ks2d.add_noise(U, 0.05)
    fp, hyper = golden_configs["full_precision"][tag], golden_configs[tag]["hyper"]
    best = [r for r in fp["table"] if r["alpha"] == hyper["alpha"] and r["threshold"] == hyper["threshold"]][0]
    return Uo, dx, dy, DT, fp["names"], np.array(best["coeffs"])
