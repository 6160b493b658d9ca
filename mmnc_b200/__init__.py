"""Importable alias of the package directory `multi-modal-neural-compression_b200/` (whose name, fixed by the
project layout, is not a valid Python identifier).  `import mmnc_b200 as mm` gives that package object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("multi-modal-neural-compression_b200")
sys.modules[__name__] = _pkg
