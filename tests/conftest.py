import importlib
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def pkg():
    """The product package (its name starts with a digit -> importlib) with the drop-in path installed."""
    mod = importlib.import_module("3dioumatch_b200")
    if not os.path.exists(mod.LIB_PATH):
        mod.build()
    mod.install_dropin()
    return mod


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as _o
    _o.build()
    return _o


@pytest.fixture(scope="session")
def ref_ext():
    """The UNMODIFIED reference extensions built into oracle/_ref (None when absent)."""
    import importlib.util
    out = {}
    for name, rel in (("_ext", "pointnet2/_ext.so"), ("iou3d_nms_cuda", "pcdet/ops/iou3d_nms/iou3d_nms_cuda.so")):
        path = os.path.join(ROOT, "oracle", "_ref", rel)
        if not os.path.exists(path):
            return None
        import torch  # noqa: F401  (libtorch must be loaded first)
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        out[name] = mod
    return out
