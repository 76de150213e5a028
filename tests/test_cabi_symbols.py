"""CPU: libb200pc.so loads and exports every symbol declared in include/*.h (no compute calls)."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names += re.findall(r"\b(b200\w*)\s*\(", text)
    return sorted(set(names))


def test_headers_declare_the_reference_surface():
    syms = declared_symbols()
    # one entry per function of pointnet2._ext (bindings.cpp:11-24) and iou3d_nms_cuda (iou3d_nms_api.cpp:11-17)
    for want in ("b200pn2_gather_points", "b200pn2_gather_points_grad", "b200pn2_furthest_point_sampling",
                 "b200pn2_three_nn", "b200pn2_three_interpolate", "b200pn2_three_interpolate_grad",
                 "b200pn2_ball_query", "b200pn2_group_points", "b200pn2_group_points_grad", "b200pn2_sa_forward",
                 "b200iou_boxes_overlap_bev", "b200iou_boxes_iou_bev", "b200iou_nms", "b200iou_boxes_iou_bev_cpu",
                 "b200iou_boxes_iou3d", "b200iou_nms_device"):
        assert want in syms, want


def test_library_exports_every_declared_symbol(pkg):
    lib = ctypes.CDLL(pkg.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export: " + name
    cabi = pkg.cabi()
    assert sorted(cabi.PROTOTYPES) == declared_symbols()  # the ctypes binding covers the headers exactly
    assert cabi.lib().b200_abi_version() == 1


def test_argument_errors_return_status_not_exit(pkg):
    cabi = pkg.cabi()
    L = cabi.lib()
    rc = L.b200pn2_furthest_point_sampling(1, 0, 4, None, None, None, None)  # N = 0 -> rejected before any launch
    assert rc != 0 and b"furthest_point_sampling" in L.b200_last_error()
    rc = L.b200pn2_three_nn(-1, 1, 1, None, None, None, None, None)
    assert rc != 0


def test_missing_library_fails_loudly(pkg, monkeypatch):
    cabi = pkg.cabi()
    monkeypatch.setattr(cabi, "_lib", None)
    monkeypatch.setattr(cabi, "LIB_PATH", "/nonexistent/libb200pc.so")
    try:
        cabi.lib()
    except RuntimeError as e:
        assert "no CPU/PyTorch fallback" in str(e)
    else:
        raise AssertionError("expected RuntimeError")
