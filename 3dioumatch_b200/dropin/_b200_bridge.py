"""Locates the 3dioumatch_b200 package from inside the drop-in tree and exposes its C-ABI binding."""
import importlib
import os
import sys

_DROPIN = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.dirname(_DROPIN)
_ROOT = os.path.dirname(_PKG)
if _ROOT not in sys.path:
    sys.path.append(_ROOT)

cabi = importlib.import_module(os.path.basename(_PKG) + "._cabi")


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream
