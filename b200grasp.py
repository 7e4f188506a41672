"""Alias module: ``import b200grasp`` loads the package in ``deep-rl-grasping_b200/``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "deep-rl-grasping_b200")
_spec = importlib.util.spec_from_file_location("b200grasp", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["b200grasp"] = _mod
_spec.loader.exec_module(_mod)
