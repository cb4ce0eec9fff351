"""Import shim: ``import smj_b200`` loads the package in ./pim-sort-merge-join_b200/ (hyphenated directory)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "pim-sort-merge-join_b200")
_spec = importlib.util.spec_from_file_location("smj_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["smj_b200"] = _mod
_spec.loader.exec_module(_mod)
