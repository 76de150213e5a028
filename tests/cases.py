"""Seeded synthetic inputs: the generators live in the package (3dioumatch_b200/synth.py, shared with bench.py);
the tests and the golden-vector generator keep importing them under this name."""
import importlib as _il
import os as _os
import sys as _sys

_ROOT = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _ROOT not in _sys.path:
    _sys.path.insert(0, _ROOT)
_synth = _il.import_module("3dioumatch_b200.synth")
globals().update({k: v for k, v in vars(_synth).items() if not k.startswith("_")})
