"""Import shim: the package directory is named `halo2-lib_b200/` (not a valid Python identifier), so
`import halo2_lib_b200` loads it from there."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "halo2-lib_b200")
_spec = importlib.util.spec_from_file_location("halo2_lib_b200", os.path.join(_d, "__init__.py"),
                                               submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["halo2_lib_b200"] = _mod
_spec.loader.exec_module(_mod)
