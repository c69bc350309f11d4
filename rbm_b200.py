"""Import shim: ``import rbm_b200`` loads the package in ``recommender-baseline-model_b200/`` (whose directory name is
not a python identifier) and registers it under this name."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "recommender-baseline-model_b200")
_spec = importlib.util.spec_from_file_location("rbm_b200", os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rbm_b200"] = _mod
_spec.loader.exec_module(_mod)
