"""Import shim: exposes the package directory ``3d-condtional-stable-diffusion_b200/`` as module ``b200dm``."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "3d-condtional-stable-diffusion_b200")
_spec = importlib.util.spec_from_file_location("b200dm", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200dm"] = _mod
_spec.loader.exec_module(_mod)
