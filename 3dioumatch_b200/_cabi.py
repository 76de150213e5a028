"""ctypes binding of libb200pc.so (include/b200_pointnet2.h, include/b200_iou3d.h, include/b200_nms.h).

There is no CPU or PyTorch fallback: if the library is missing, every operator raises."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200_LIB_PATH") or os.path.join(_HERE, "lib", "libb200pc.so")  # override: developer builds

c_int, c_float, c_void_p, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t


class MlpLayer(ctypes.Structure):
    """b200_mlp_layer"""
    _fields_ = [("cin", c_int), ("cout", c_int), ("weight", c_void_p), ("scale", c_void_p), ("shift", c_void_p)]


class BnLayer(ctypes.Structure):
    """b200_bn_layer"""
    _fields_ = [("cin", c_int), ("cout", c_int), ("weight", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
                ("running_mean", c_void_p), ("running_var", c_void_p)]


# name -> (restype, argtypes); every symbol declared in include/*.h
PROTOTYPES = {
    "b200_abi_version": (c_int, []),
    "b200_last_error": (ctypes.c_char_p, []),
    "b200_launch_count": (ctypes.c_ulonglong, []),
    "b200pn2_fps_set_policy": (c_int, [c_int]),
    "b200pn2_fps_force_shape": (c_int, [c_int, c_int, c_int]),
    "b200pn2_fps_set_prefix_speculation": (c_int, [c_int]),
    "b200pn2_furthest_point_sampling": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200pn2_gather_points": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200pn2_gather_points_grad": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200pn2_ball_query": (c_int, [c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200pn2_group_points": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200pn2_group_points_grad": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200pn2_three_nn": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200pn2_three_interpolate": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200pn2_three_interpolate_grad": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                c_void_p]),
    "b200pn2_scatter_det_workspace": (c_size_t, [c_int, c_int, c_int]),
    "b200pn2_gather_points_grad_det": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_size_t, c_void_p]),
    "b200pn2_group_points_grad_det": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_size_t, c_void_p]),
    "b200pn2_three_interpolate_grad_det": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_size_t, c_void_p]),
    "b200pn2_sa_forward_workspace": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "b200pn2_sa_forward": (c_int, [c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(MlpLayer), c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200pn2_interp_mlp_forward": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, ctypes.POINTER(MlpLayer), c_void_p, c_void_p, c_size_t,
                                           c_void_p]),
    "b200pn2_mlp_plan_bytes": (c_size_t, [c_int, c_int, c_int, ctypes.POINTER(MlpLayer), c_int, c_int]),
    "b200pn2_mlp_plan_build": (c_int, [c_int, c_int, c_int, ctypes.POINTER(MlpLayer), c_int, c_int, c_void_p, c_size_t,
                                       c_void_p]),
    "b200pn2_sa_forward_planned": (c_int, [c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(MlpLayer), c_void_p,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "b200pn2_interp_mlp_forward_planned": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_void_p, c_int, ctypes.POINTER(MlpLayer), c_void_p,
                                                   c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "b200pn2_fp_rows_forward": (c_int, [c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                        ctypes.POINTER(MlpLayer), c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                        c_void_p]),
    "b200pn2_row_mlp_forward": (c_int, [c_int, c_int, c_int, c_int, c_void_p, c_int, ctypes.POINTER(MlpLayer), c_int,
                                        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200pn2_row_mlp_forward_cm": (c_int, [c_int, c_int, c_int, c_void_p, c_int, ctypes.POINTER(MlpLayer), c_int,
                                           c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "b200pn2_transpose_cn": (c_int, [c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "b200pn2_sa_train_saved_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(BnLayer), c_int]),
    "b200pn2_sa_train_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, ctypes.POINTER(BnLayer), c_int,
                                                    c_int]),
    "b200pn2_sa_train_forward": (c_int, [c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p, c_int, ctypes.POINTER(BnLayer), c_float, c_float,
                                         c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_void_p]),
    "b200pn2_sa_train_backward": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                          ctypes.POINTER(BnLayer), c_void_p, c_void_p, c_size_t, c_void_p, c_void_p,
                                          ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p), ctypes.POINTER(c_void_p),
                                          c_void_p, c_size_t, c_void_p]),
    "b200pn2_sa_tensor_work": (c_int, [ctypes.POINTER(ctypes.c_ulonglong), c_int]),
    "b200iou_boxes_overlap_bev": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200iou_boxes_iou_bev": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200iou_boxes_iou3d": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200iou_boxes_iou3d_batched": (c_int, [c_int, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "b200iou_nms_device": (c_int, [c_int, c_void_p, c_float, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "b200iou_nms": (c_int, [c_int, c_void_p, c_float, c_int, c_void_p, ctypes.POINTER(c_int), c_void_p]),
    "b200iou_boxes_iou_bev_cpu": (c_int, [c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "b200nms_aabb_suppress": (c_int, [c_int, c_int, c_int, c_int, c_int, ctypes.c_double, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_void_p]),
    "b200nms_box_extents": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """Load libb200pc.so and bind every prototype.  Raises RuntimeError if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "3dioumatch_b200: %s is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C 3dioumatch_b200/csrc`; there is no CPU/PyTorch fallback." % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)  # AttributeError here == header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().b200_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg.decode() if msg else "?"))


def launch_count():
    return int(lib().b200_launch_count())


def sa_tensor_work(reset=True):
    """Executed TF32 FLOPs of the fused tensor-core kernels on the current device since the last reset (synchronises)."""
    n = ctypes.c_ulonglong(0)
    check(lib().b200pn2_sa_tensor_work(ctypes.byref(n), 1 if reset else 0), "sa_tensor_work")
    return int(n.value) * 2048


def set_fps_policy(policy):
    """'latency' (default; shortest serial chain) or 'throughput' (least SM-time; pipelined steps).  Returns the previous
    policy name.  Results are bit-identical either way (include/b200_pointnet2.h: b200pn2_fps_set_policy)."""
    names = ("latency", "throughput")
    prev = lib().b200pn2_fps_set_policy(names.index(policy))
    return names[prev]


def force_fps_shape(kernel=-1, cluster=0, threads=0):
    """Tuning / test hook: kernel 0 fps_owner_kernel wherever it applies, 2 fps_cluster_kernel (round 1), -1 default;
    cluster / threads 0 = cost model (include/b200_pointnet2.h: b200pn2_fps_force_shape)."""
    lib().b200pn2_fps_force_shape(int(kernel), int(cluster), int(threads))


def set_fps_prefix_speculation(mode):
    """1 on, 0 off, -1 default (on): verify-then-skip for inputs already in furthest-point order (the hierarchical levels);
    returns the previous mode (include/b200_pointnet2.h: b200pn2_fps_set_prefix_speculation)."""
    return int(lib().b200pn2_fps_set_prefix_speculation(int(mode)))
