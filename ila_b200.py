"""Import alias: the package directory `infinitelinearalgebra.jl_b200/` has a dot in its name, which Python's
import statement cannot spell.  `import ila_b200` loads that directory as the package `ila_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "infinitelinearalgebra.jl_b200")
_spec = importlib.util.spec_from_file_location("ila_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ila_b200"] = _mod
_spec.loader.exec_module(_mod)
