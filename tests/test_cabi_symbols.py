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


def test_shipped_library_holds_the_blackwell_native_instructions(pkg):
    """The product library is sm_100a code built around the instructions DESIGN.md names, not a recompiled generic path:
    tcgen05.mma / commit / ld / st (UTCHMMA, UTCBAR, LDTM, STTM), cp.async.bulk (UBLKCP), mbarrier (SYNCS), redux.sync
    (CREDUX), cluster barriers and DSMEM st.async (UCGABAR_ARV, STAS), packed fp32 pairs (FFMA2).  SASS is read with
    cuobjdump from the very file the GPU tests load; skipped where the CUDA toolkit is not installed."""
    import shutil
    import subprocess

    import pytest
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not installed")
    elf = subprocess.run([tool, "-lelf", pkg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in elf, elf[:200]
    sass = subprocess.run([tool, "-sass", pkg.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic, at_least in (("UTCHMMA", 100), ("UTCBAR", 20), ("LDTM", 10), ("STTM", 5), ("UBLKCP", 5), ("SYNCS", 200),
                               ("CREDUX", 50), ("UCGABAR_ARV", 10), ("STAS", 10), ("FFMA2", 100)):
        n = len(re.findall(r"\b%s\b" % mnemonic, sass))
        assert n >= at_least, "%s: %d occurrences" % (mnemonic, n)
    # the tensor-core kernel variants the launcher can select, and both FPS generations, are all in the image
    for kernel in ("sa_tcp_kernel", "fps_owner_kernel", "fps_cluster_kernel", "dw_tc_kernel", "bg_query_kernel",
                   "pair_kernel", "nms_mask_kernel", "three_nn_kernel"):
        assert kernel in sass, kernel


def _header_prototypes():
    """name -> (return type text, [parameter type texts]) parsed from include/*.h (comments stripped)."""
    protos = {}
    for h in sorted(glob.glob(os.path.join(ROOT, "include", "*.h"))):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        text = re.sub(r"//[^\n]*", "", text)
        for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(b200\w*)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
            ret, name, params = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
            plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
            protos[name] = (ret, plist)
    return protos


def _kind(ctype_text):
    """Coarse ABI class of a C parameter / return type."""
    t = ctype_text.replace("const", " ").strip()
    if "*" in t or "b200_stream_t" in t:
        return "ptr"
    if "size_t" in t or "unsigned long long" in t:
        return "u64"            # both 8-byte unsigned on the LP64 targets this library is built for
    if "double" in t:
        return "double"
    if "float" in t:
        return "float"
    if re.search(r"\bint\b|int32_t", t):
        return "int"
    raise AssertionError("unclassified C type: %r" % ctype_text)


def test_ctypes_signatures_match_the_headers(pkg):
    """Every prototype of include/*.h against the ctypes table: same number of parameters, and each parameter of the same
    ABI class (int / float / 8-byte unsigned / pointer) in the same position -- an argument dropped or reordered on either side
    would still load and then read garbage."""
    import ctypes
    cabi = pkg.cabi()
    protos = _header_prototypes()
    assert sorted(protos) == sorted(cabi.PROTOTYPES)

    def ckind(t):
        if t is ctypes.c_int:
            return "int"
        if t is ctypes.c_float:
            return "float"
        if t is ctypes.c_double:
            return "double"
        if t in (ctypes.c_size_t, ctypes.c_ulonglong):
            return "u64"
        if t in (ctypes.c_void_p, ctypes.c_char_p) or (isinstance(t, type) and issubclass(t, ctypes._Pointer)):
            return "ptr"
        raise AssertionError("unclassified ctypes type: %r" % (t,))

    for name, (ret, params) in sorted(protos.items()):
        restype, argtypes = cabi.PROTOTYPES[name]
        assert len(params) == len(argtypes), "%s: header has %d parameters, ctypes %d" % (name, len(params), len(argtypes))
        got = [ckind(t) for t in argtypes]
        want = [_kind(p) for p in params]
        assert got == want, "%s: parameter classes differ\n  header %s\n  ctypes %s" % (name, want, got)
        assert ckind(restype) == _kind(ret), "%s: return type %r vs %r" % (name, ret, restype)
