"""Import alias: the package directory is `show-o_b200/` (not a valid identifier), so `import showo_b200` loads it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "show-o_b200")
_spec = importlib.util.spec_from_file_location("showo_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["showo_b200"] = _mod
_spec.loader.exec_module(_mod)
