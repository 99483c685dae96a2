"""Loader for the UNMODIFIED reference scripts staged under ``baseline/_ref`` (git-ignored).

TEST / BASELINE INFRASTRUCTURE ONLY (same rule as the rest of ``oracle/``): ``tests/``,
``__graft_entry__`` and ``bench.py``'s CPU legs may use it; the product never does.

``/root/reference`` only exists in the build container.  ``__graft_entry__.build()`` copies its
``scripts/`` and ``examples/`` directories verbatim into ``baseline/_ref/`` (never into git:
``.gitignore`` lists the directory; ``gpurun`` ships it to the GPU box), so that on the box
(a) the drop-in can be exercised on the real ``main()`` functions and (b) the reference's own
CPU path can be timed beside the GPU path (``cpu_baseline.kind = "reference"``).

The scripts import matplotlib (absent from the image: stubbed with MagicMock, as in
``tests/golden/make_golden.py``) and ``patch_based_pde_discovery.py`` creates its output
directory at import time (redirected to a temporary directory through its ``PDE_OUTPUT_DIR``-free
layout by suppressing ``Path.mkdir`` during the import).
"""

from __future__ import annotations

import importlib.util
import shutil
import sys
from pathlib import Path
from unittest import mock

ROOT = Path(__file__).resolve().parent.parent
REF_SRC = Path("/root/reference")
REF_DIR = ROOT / "baseline" / "_ref"

FILES = {"ks2d": ("scripts", "ks2d_stridge_benchmark.py"), "basic": ("examples", "basic_usage.py"),
         "patch": ("scripts", "patch_based_pde_discovery.py"), "sindy": ("scripts", "patch_based_sindy.py"),
         "analyze": ("scripts", "analyze_results.py")}


def stage() -> bool:
    """Copy the reference's scripts/ and examples/ into baseline/_ref (build container only)."""
    if not REF_SRC.exists():
        return REF_DIR.exists()
    for d in ("scripts", "examples"):
        shutil.copytree(REF_SRC / d, REF_DIR / d, dirs_exist_ok=True)
    return True


def available(which: str = "ks2d") -> bool:
    sub, name = FILES[which]
    return (REF_DIR / sub / name).exists()


def path_of(which: str) -> Path:
    sub, name = FILES[which]
    return REF_DIR / sub / name


def load(which: str, fresh: bool = False):
    """Import one reference script as a module (``ref_<which>``); ``fresh`` re-executes it."""
    modname = f"ref_{which}"
    if not fresh and modname in sys.modules:
        return sys.modules[modname]
    if which == "analyze":
        raise RuntimeError("analyze_results.py is module-level code that needs data/Real-Images; use source_lines()")
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.gridspec", "matplotlib.patches", "matplotlib.colors"):
        sys.modules.setdefault(m, mock.MagicMock())
    spec = importlib.util.spec_from_file_location(modname, path_of(which))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    with mock.patch.object(Path, "mkdir", lambda *a, **k: None):
        spec.loader.exec_module(mod)
    return mod


def source_lines(which: str, first: int, last: int) -> str:
    """Lines first..last (1-based, inclusive) of a staged reference script, for scripts that cannot be imported."""
    text = path_of(which).read_text().splitlines()
    return "\n".join(text[first - 1:last]) + "\n"
