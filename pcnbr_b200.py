"""`import pcnbr_b200` -- importable alias of the package directory `3d-semantic-segmentation-benchmark_b200/`
(whose name is not a Python identifier).  Put the repository root on sys.path / PYTHONPATH; this module replaces itself
in sys.modules by the real package, so `pcnbr_b200.ops`, `from pcnbr_b200.common import SetAbstraction` ... work as
for any package."""
import importlib.util
import os
import sys

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3d-semantic-segmentation-benchmark_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_DIR, "__init__.py"), submodule_search_locations=[_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
try:
    _spec.loader.exec_module(_mod)
except BaseException:
    sys.modules.pop(__name__, None)
    raise
