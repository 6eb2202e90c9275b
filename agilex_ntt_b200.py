"""Import alias: the package directory is `agilex-ntt_b200/` (not a valid Python identifier), so this module loads
it under the name `agilex_ntt_b200` and replaces itself in sys.modules."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "agilex-ntt_b200")
_spec = importlib.util.spec_from_file_location("agilex_ntt_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["agilex_ntt_b200"] = _mod
_spec.loader.exec_module(_mod)
