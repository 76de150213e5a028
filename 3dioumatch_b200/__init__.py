"""3dioumatch_b200 -- B200-native (sm_100a) PointNet++ set-abstraction and rotated 3D-IoU/NMS operators
behind the operator surface of yezhen17/3DIoUMatch.

Layout
  csrc/      hand-written CUDA kernels + the C ABI (include/b200_pointnet2.h, include/b200_iou3d.h)
  lib/       libb200pc.so (built in-tree by `make -C 3dioumatch_b200/csrc` or __graft_entry__.build())
  _cabi.py   ctypes binding of the C ABI (fails loudly when the library is missing)
  dropin/    the reference's Python operator surface re-implemented on top of the C ABI:
               pointnet2/{_ext,pointnet2_utils,pointnet2_modules,pytorch_utils}.py
               pcdet/ops/iou3d_nms/{iou3d_nms_cuda,iou3d_nms_utils}.py

The package name starts with a digit, so it is imported with importlib:
    pkg = importlib.import_module("3dioumatch_b200"); pkg.install_dropin()
    import pointnet2.pointnet2_utils, pointnet2_modules          # as the reference's callers do
    from pcdet.ops.iou3d_nms import iou3d_nms_utils
"""
import os
import subprocess
import sys

__version__ = "0.1.0"

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
DROPIN_DIR = os.path.join(PKG_DIR, "dropin")
LIB_PATH = os.path.join(PKG_DIR, "lib", "libb200pc.so")


def build(verbose=False):
    """Compile libb200pc.so for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(PKG_DIR, "csrc"), "-j8"]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout)
    if out.returncode != 0:
        raise RuntimeError("building libb200pc.so failed")
    return LIB_PATH


def install_dropin():
    """Put the drop-in `pointnet2` / `pcdet` packages (and the reference-style flat module path
    `pointnet2_modules`, used by models/backbone_module.py:16-19) at the front of sys.path."""
    for p in (os.path.join(DROPIN_DIR, "pointnet2"), DROPIN_DIR):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    return DROPIN_DIR


def cabi():
    from . import _cabi
    return _cabi
