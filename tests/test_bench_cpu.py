"""CPU checks of bench.py's reference arm (the only leg that may execute oracle/ outside tests):
the port runs on a tiny sample, recovers the same model as the single-process oracle, and the
printed line carries the contract keys."""

import importlib.util
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _bench():
    spec = importlib.util.spec_from_file_location("bench", ROOT / "bench.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["bench"] = mod  # multiprocessing pickles the worker function by module name
    spec.loader.exec_module(mod)
    return mod


def test_cpu_port_matches_single_process_oracle():
    import multiprocessing as mp

    from oracle import gram, ks2d as O

    b = _bench()
    U = b.cpu_sample(frames=13, size=32)
    b._CPU["U"] = U
    with mp.get_context("fork").Pool(2) as pool:
        best = b.cpu_port_step(U, 2, pool)
    names, terms = O.build_dictionary_true(U[:-1], b.D0, b.D1)
    X, y = O.build_blockwise_dataset((U[1:] - U[:-1]) / b.DT, terms, names, block_t=3, block_x=8, block_y=8)
    rows_per_tb = len(y) // 4
    s_tr, s_te = gram.pack_stats(X[: 2 * rows_per_tb], y[: 2 * rows_per_tb]), gram.pack_stats(X[2 * rows_per_tb:], y[2 * rows_per_tb:])
    ref = gram.ks_fit_from_stats(s_tr, s_te, 3, alphas=O.GRID_ALPHAS, thresholds=O.GRID_THRESHOLDS)
    assert (best["alpha"], best["threshold"]) == (ref["alpha"], ref["threshold"])
    np.testing.assert_allclose(best["coeffs"], ref["coeffs"], rtol=1e-9)


def _reference_arm(extra, env_extra):
    import os

    env = dict(os.environ, PG_BENCH_REF_SAMPLE="13,64", PG_BENCH_PORT_SAMPLE="13,64", **env_extra)
    out = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"] + extra,
                         cwd=ROOT, capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["impl"] == "reference" and line["dtype"] == "f64" and line["vs_baseline"] is None
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0
    assert line["e2e"]["value"] == line["value"] == line["cpu_baseline"]["value"]
    return line


def test_reference_arm_prints_contract_line_port():
    line = _reference_arm(["--port-only"], {})
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1


def test_reference_arm_times_the_staged_reference():
    """With baseline/_ref staged (the build container, the GPU box) the arm times the UNMODIFIED reference functions
    and reports the port beside it; the reference's selection equals the port's on the same sample."""
    from oracle import refload

    if not refload.available("ks2d"):
        import pytest

        pytest.skip("baseline/_ref is not staged (run __graft_entry__.build() where /root/reference exists)")
    line = _reference_arm([], {})
    assert line["cpu_baseline"]["kind"] == "reference" and line["cpu_baseline"]["cores"] == 1
    assert line["port"]["kind"] == "port" and line["port"]["value"] > 0


def test_reference_step_agrees_with_the_port():
    import multiprocessing as mp

    from oracle import refload

    if not refload.available("ks2d"):
        import pytest

        pytest.skip("baseline/_ref is not staged")
    b = _bench()
    U = b.cpu_sample(frames=13, size=32)
    ref = b.reference_step(refload.load("ks2d"), U)
    b._CPU["U"] = U
    with mp.get_context("fork").Pool(2) as pool:
        port = b.cpu_port_step(U, 4, pool)       # 4 t-blocks -> the same 2 / 2 train / test split as reference_step
    assert (ref["alpha"], ref["threshold"]) == (port["alpha"], port["threshold"])
    np.testing.assert_allclose(ref["coeffs"], port["coeffs"], rtol=1e-8)
