"""Import shim: the product package lives in ``pde-discovery-laser-matter_b200/`` (the name the
build contract fixes, which is not a valid Python identifier).  Pointing ``__path__`` there
makes ``import pde_b200`` / ``pde_b200.ks2d`` resolve to that directory."""

from pathlib import Path

_real = Path(__file__).resolve().parent.parent / "pde-discovery-laser-matter_b200"
__path__ = [str(_real)]
exec(compile((_real / "__init__.py").read_text(), str(_real / "__init__.py"), "exec"))
