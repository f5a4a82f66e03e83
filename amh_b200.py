"""Loader for the package directory `advancedmh.jl_b200/` (its name contains a dot, so it is
imported under the module name `advancedmh_jl_b200`).  `import amh_b200 as amh` gives the package."""
import importlib.util
import os
import sys

_NAME = "advancedmh_jl_b200"
_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "advancedmh.jl_b200")

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_DIR, "__init__.py"),
                                                   submodule_search_locations=[_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)

_pkg = sys.modules[_NAME]
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
package = _pkg
