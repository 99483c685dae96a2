"""pde_b200 -- B200-native (sm_100a) hot path of anpeata/pde-discovery-laser-matter.

The directory is named ``pde-discovery-laser-matter_b200`` (not an importable identifier);
import it as ``pde_b200`` (the top-level shim package points its ``__path__`` here).

    pde_b200.ks2d         drop-in for scripts/ks2d_stridge_benchmark.py hot-path functions
    pde_b200.basic_usage  drop-in for examples/basic_usage.py
    pde_b200.patch        drop-in for scripts/patch_based_pde_discovery.py
    pde_b200.analyze      the analyze_results.py dialect (slice-central FD, Models 1-6, rollout checks)
    pde_b200.sindy        scripts/patch_based_sindy.py (periodic-roll patches, 11-term library, weighted ensemble)
    pde_b200.ops          one function per C-ABI entry point (include/pdegram.h)
    pde_b200.slabs        time-slab sharding across GPUs (halo frame + Gram all-reduce)
    pde_b200.patch_reference(module)   rebind a loaded reference script to the GPU functions

Everything computes through libpdegram.so (hand-written CUDA); there is no CPU fallback.
"""

from . import _lib
from ._lib import PdeGramError, build, load
from .drop_in import patch_reference
from . import analyze, basic_usage, ks2d, ops, patch, sindy  # noqa: E402  (torch is imported lazily, on first compute call)

__all__ = ["PdeGramError", "build", "load", "patch_reference", "ks2d", "basic_usage", "patch", "analyze", "sindy", "ops"]
__version__ = "0.1.0"
