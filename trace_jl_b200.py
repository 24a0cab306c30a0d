"""Import shim: the package directory is named `trace.jl_b200/` (as the layout contract asks), which is not an
importable identifier.  `import trace_jl_b200` loads that directory as the package `trace_jl_b200`."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "trace.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "trace_jl_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["trace_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
