"""Importable alias of the ``3dhandposeestimation_b200`` package (whose directory name is
not a valid Python identifier): ``import handpose_b200 as hp; hp.ManoLayer(...)``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("3dhandposeestimation_b200")
sys.modules[__name__] = _pkg
