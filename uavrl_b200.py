"""Import shim: makes the package directory `dqn-based-uav-3d_path_planer_b200/` importable as
`uavrl_b200` (hyphens are not legal in Python module names)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dqn-based-uav-3d_path_planer_b200")
_spec = importlib.util.spec_from_file_location("uavrl_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["uavrl_b200"] = _mod
_spec.loader.exec_module(_mod)
