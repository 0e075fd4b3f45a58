"""Import shim: `import mmt_psm_b200` loads the package that lives in the directory `mmt-psm_b200/`
(a hyphen is not a legal Python identifier, the directory name is fixed by the project layout)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mmt-psm_b200")
_spec = importlib.util.spec_from_file_location(
    "mmt_psm_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mmt_psm_b200"] = _mod
_spec.loader.exec_module(_mod)
