"""`import gaitk` -> the package directory (its name is not a Python identifier)."""
import importlib
import sys
from pathlib import Path

_root = str(Path(__file__).resolve().parent)
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("towards-relaxed-multimodal-inputs-for-gait-based-parkinson-s-disease-assessment_b200")
sys.modules[__name__] = _pkg
