"""Import shim: the package directory is `stochastic-epidemic-modelling_b200/` (hyphens are not valid in a
Python identifier), so `import sem_b200` loads it under that name."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stochastic-epidemic-modelling_b200")
_spec = importlib.util.spec_from_file_location("sem_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["sem_b200"] = _mod
_spec.loader.exec_module(_mod)
