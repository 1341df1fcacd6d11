"""Import shim: ``import owc_b200`` loads the package that lives in ``lmms-owc_b200/``.

The package directory is named after the reference (``lmms-owc`` + ``_b200``) and a hyphen is not
a legal Python identifier, so this one-file loader registers that directory under the importable
name ``owc_b200`` (sub-modules resolve through ``submodule_search_locations``).
"""

import importlib.util as _ilu
import os as _os
import sys as _sys

_pkg_dir = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "lmms-owc_b200")
_spec = _ilu.spec_from_file_location(
    "owc_b200", _os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir]
)
_mod = _ilu.module_from_spec(_spec)
_sys.modules["owc_b200"] = _mod
_spec.loader.exec_module(_mod)
