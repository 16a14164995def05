"""Import shim: the package lives in `vec-ode_b200/`, whose name is not a Python identifier."""
import importlib.util as _u
import os as _os
import sys as _sys

_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "vec-ode_b200")
_spec = _u.spec_from_file_location("vecode_b200", _os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = _u.module_from_spec(_spec)
_sys.modules["vecode_b200"] = _mod
_spec.loader.exec_module(_mod)
