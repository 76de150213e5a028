"""Drop-in for the reference's native module `pointnet2._ext`
(pointnet2/_ext_src/src/bindings.cpp:11-24): same nine functions, same argument order, dtypes, shapes and
error behaviour (RuntimeError for non-contiguous / wrong dtype / CPU tensors, utils.h:10-30), implemented
by the hand-written sm_100a kernels of libb200pc.so through its C ABI.  Outputs are allocated by the callee
on the input's device; launches go to the current torch stream.  No CPU path exists ("CPU not supported",
as in the reference)."""
import ctypes

import torch

from _b200_bridge import cabi, stream_ptr

_L = cabi.lib


def _chk(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _contig(t, name):
    _chk(t.is_contiguous(), "%s must be a contiguous tensor" % name)


def _is_float(t, name):
    _chk(t.dtype == torch.float32, "%s must be a float tensor" % name)


def _is_int(t, name):
    _chk(t.dtype == torch.int32, "%s must be an int tensor" % name)


def _cuda(t, name):
    _chk(t.is_cuda, "CPU not supported" if name is None else "%s must be a CUDA tensor" % name)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None and t.numel() > 0 else ctypes.c_void_p(0)


def deterministic():
    """True when gradients must be bit-reproducible: torch.use_deterministic_algorithms(True) or B200_DETERMINISTIC=1.
    The *_grad entries then run the sorted-segment kernels (csrc/scatter_det.cu) instead of the reference's atomicAdd
    scatter (sampling_gpu.cu:39-52, group_points_gpu.cu:48-68, interpolate_gpu.cu:121-148)."""
    import os
    return torch.are_deterministic_algorithms_enabled() or os.environ.get("B200_DETERMINISTIC") == "1"


def _det_workspace(B, targets, entries, device):
    nbytes = int(_L().b200pn2_scatter_det_workspace(int(B), int(targets), int(entries)))
    return torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=device), nbytes


def furthest_point_sampling(points, nsamples):
    """(B,N,3) f32 -> (B,nsamples) int32   [sampling.cpp:70-91]"""
    _contig(points, "points"); _is_float(points, "points"); _cuda(points, None)
    B, N = points.size(0), points.size(1)
    out = torch.empty((B, int(nsamples)), dtype=torch.int32, device=points.device)
    if B == 0 or nsamples == 0:
        return out
    with torch.cuda.device(points.device):
        cabi.check(_L().b200pn2_furthest_point_sampling(B, N, int(nsamples), _p(points), _p(out), None, stream_ptr()),
                   "furthest_point_sampling")
    return out


def gather_points(points, idx):
    """(B,C,N) f32, (B,m) i32 -> (B,C,m)   [sampling.cpp:20-44]"""
    _contig(points, "points"); _contig(idx, "idx"); _is_float(points, "points"); _is_int(idx, "idx")
    _cuda(points, None); _cuda(idx, "idx")
    B, C, N = points.shape
    m = idx.size(1)
    out = torch.empty((B, C, m), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        cabi.check(_L().b200pn2_gather_points(B, C, N, m, _p(points), _p(idx), _p(out), stream_ptr()), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """(B,C,m) f32, (B,m) i32, n -> (B,C,n)   [sampling.cpp:46-69]"""
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _is_float(grad_out, "grad_out"); _is_int(idx, "idx")
    _cuda(grad_out, None); _cuda(idx, "idx")
    B, C, m = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        if deterministic():
            ws, nbytes = _det_workspace(B, n, m, grad_out.device)
            cabi.check(_L().b200pn2_gather_points_grad_det(B, C, int(n), m, _p(grad_out), _p(idx), _p(out), _p(ws), nbytes,
                                                           stream_ptr()), "gather_points_grad_det")
        else:
            cabi.check(_L().b200pn2_gather_points_grad(B, C, int(n), m, _p(grad_out), _p(idx), _p(out), stream_ptr()),
                       "gather_points_grad")
    return out


def three_nn(unknowns, knows):
    """(B,n,3), (B,m,3) -> [dist2 (B,n,3) f32, idx (B,n,3) i32]   [interpolate.cpp:19-45]"""
    _contig(unknowns, "unknowns"); _contig(knows, "knows"); _is_float(unknowns, "unknowns"); _is_float(knows, "knows")
    _cuda(unknowns, None); _cuda(knows, "knows")
    B, n = unknowns.size(0), unknowns.size(1)
    m = knows.size(1)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    with torch.cuda.device(unknowns.device):
        cabi.check(_L().b200pn2_three_nn(B, n, m, _p(unknowns), _p(knows), _p(dist2), _p(idx), stream_ptr()), "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """(B,c,m) f32, (B,n,3) i32, (B,n,3) f32 -> (B,c,n)   [interpolate.cpp:47-74]"""
    for t, nm in ((points, "points"), (idx, "idx"), (weight, "weight")):
        _contig(t, nm)
    _is_float(points, "points"); _is_int(idx, "idx"); _is_float(weight, "weight")
    _cuda(points, None); _cuda(idx, "idx"); _cuda(weight, "weight")
    B, C, m = points.shape
    n = idx.size(1)
    out = torch.empty((B, C, n), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        cabi.check(_L().b200pn2_three_interpolate(B, C, m, n, _p(points), _p(idx), _p(weight), _p(out), stream_ptr()),
                   "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """(B,c,n) f32, idx, weight, m -> (B,c,m): the true scatter-add gradient (interpolate_gpu.cu:121-148).
    The reference's host wrapper launches the forward kernel here by mistake (interpolate.cpp:95)."""
    for t, nm in ((grad_out, "grad_out"), (idx, "idx"), (weight, "weight")):
        _contig(t, nm)
    _is_float(grad_out, "grad_out"); _is_int(idx, "idx"); _is_float(weight, "weight")
    _cuda(grad_out, None); _cuda(idx, "idx"); _cuda(weight, "weight")
    B, C, n = grad_out.shape
    out = torch.empty((B, C, int(m)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        if deterministic():
            ws, nbytes = _det_workspace(B, m, 3 * n, grad_out.device)
            cabi.check(_L().b200pn2_three_interpolate_grad_det(B, C, n, int(m), _p(grad_out), _p(idx), _p(weight), _p(out),
                                                               _p(ws), nbytes, stream_ptr()), "three_interpolate_grad_det")
        else:
            cabi.check(_L().b200pn2_three_interpolate_grad(B, C, n, int(m), _p(grad_out), _p(idx), _p(weight), _p(out),
                                                           stream_ptr()), "three_interpolate_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """(B,M,3), (B,N,3), radius, nsample -> (B,M,nsample) i32   [ball_query.cpp:13-37]"""
    _contig(new_xyz, "new_xyz"); _contig(xyz, "xyz"); _is_float(new_xyz, "new_xyz"); _is_float(xyz, "xyz")
    _cuda(new_xyz, None); _cuda(xyz, "xyz")
    B, M = new_xyz.size(0), new_xyz.size(1)
    N = xyz.size(1)
    idx = torch.empty((B, M, int(nsample)), dtype=torch.int32, device=new_xyz.device)
    with torch.cuda.device(new_xyz.device):
        cabi.check(_L().b200pn2_ball_query(B, N, M, float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx),
                                           stream_ptr()), "ball_query")
    return idx


def group_points(points, idx):
    """(B,C,N) f32, (B,M,ns) i32 -> (B,C,M,ns)   [group_points.cpp:17-39]"""
    _contig(points, "points"); _contig(idx, "idx"); _is_float(points, "points"); _is_int(idx, "idx")
    _cuda(points, None); _cuda(idx, "idx")
    B, C, N = points.shape
    M, ns = idx.size(1), idx.size(2)
    out = torch.empty((B, C, M, ns), dtype=torch.float32, device=points.device)
    with torch.cuda.device(points.device):
        cabi.check(_L().b200pn2_group_points(B, C, N, M, ns, _p(points), _p(idx), _p(out), stream_ptr()), "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """(B,C,M,ns) f32, idx, n -> (B,C,n)   [group_points.cpp:41-65]"""
    _contig(grad_out, "grad_out"); _contig(idx, "idx"); _is_float(grad_out, "grad_out"); _is_int(idx, "idx")
    _cuda(grad_out, None); _cuda(idx, "idx")
    B, C, M, ns = grad_out.shape
    out = torch.empty((B, C, int(n)), dtype=torch.float32, device=grad_out.device)
    with torch.cuda.device(grad_out.device):
        if deterministic():
            ws, nbytes = _det_workspace(B, n, M * ns, grad_out.device)
            cabi.check(_L().b200pn2_group_points_grad_det(B, C, int(n), M, ns, _p(grad_out), _p(idx), _p(out), _p(ws), nbytes,
                                                          stream_ptr()), "group_points_grad_det")
        else:
            cabi.check(_L().b200pn2_group_points_grad(B, C, int(n), M, ns, _p(grad_out), _p(idx), _p(out), stream_ptr()),
                       "group_points_grad")
    return out


# ---- wide entries (no reference counterpart): fused stages on the tensor-core SharedMLP kernels ----------------------

def _layer_array(layers):
    """[(weight (cout,cin), scale (cout,), shift (cout,))] -> (b200_mlp_layer[], tensors kept alive by the caller)."""
    arr = (cabi.MlpLayer * len(layers))()
    keep = []
    for i, (w, sc, sh) in enumerate(layers):
        for t, nm in ((w, "weight"), (sc, "scale"), (sh, "shift")):
            _contig(t, nm); _is_float(t, nm); _cuda(t, nm)
        w2 = w.reshape(w.size(0), -1)
        keep.append(w2)
        arr[i].cin, arr[i].cout = w2.size(1), w2.size(0)
        arr[i].weight, arr[i].scale, arr[i].shift = w2.data_ptr(), sc.data_ptr(), sh.data_ptr()
    return arr, keep


def _plan_args(plan):
    if plan is None:
        return ctypes.c_void_p(0), 0
    _chk(plan.is_cuda and plan.dtype == torch.uint8 and plan.is_contiguous(), "plan must be a contiguous CUDA uint8 tensor")
    return _p(plan), plan.numel()


def mlp_plan(layers, C_feat, use_xyz, row_output=False, plain_rows=False):
    """Pack the weights of a SharedMLP stack once (include/b200_pointnet2.h: b200pn2_mlp_plan_build).  Returns a CUDA
    uint8 tensor to pass as `plan=` to the fused entries, or None when the tensor-core kernel does not take the stack.
    Valid until the weights / folded affine change."""
    arr, keep = _layer_array(layers)
    dev = layers[0][0].device
    with torch.cuda.device(dev):
        n = int(_L().b200pn2_mlp_plan_bytes(int(C_feat), int(bool(use_xyz)), len(layers), arr, int(bool(row_output)),
                                            int(bool(plain_rows))))
        if n == 0:
            return None
        plan = torch.empty((n,), dtype=torch.uint8, device=dev)
        cabi.check(_L().b200pn2_mlp_plan_build(int(C_feat), int(bool(use_xyz)), len(layers), arr, int(bool(row_output)),
                                               int(bool(plain_rows)), _p(plan), n, stream_ptr()), "mlp_plan_build")
    return plan


def sa_forward(xyz, features, new_xyz, radius, nsample, layers, use_xyz=True, normalize_xyz=False,
               features_pm=None, idx=None, want_idx=False, want_pm=False, plan=None):
    """Fused set-abstraction forward (include/b200_pointnet2.h: b200pn2_sa_forward_planned).

    xyz (B,N,3), features (B,C,N) or None, new_xyz (B,M,3); layers = [(weight (cout,cin), scale (cout,), shift (cout,))]
    returns (out (B,cout,M), out_pm (B,M,cout) or None, idx (B,M,nsample) or None)."""
    _contig(xyz, "xyz"); _contig(new_xyz, "new_xyz"); _is_float(xyz, "xyz"); _is_float(new_xyz, "new_xyz")
    _cuda(xyz, None); _cuda(new_xyz, "new_xyz")
    B, N = xyz.size(0), xyz.size(1)
    M = new_xyz.size(1)
    C = 0
    if features is not None:
        _contig(features, "features"); _is_float(features, "features"); _cuda(features, "features")
        C = features.size(1)
    if features_pm is not None:
        _contig(features_pm, "features_pm"); _is_float(features_pm, "features_pm"); _cuda(features_pm, "features_pm")
        C = features_pm.size(2)
    if idx is not None:
        _contig(idx, "idx"); _is_int(idx, "idx"); _cuda(idx, "idx")
    arr, keep_alive = _layer_array(layers)
    cout = layers[-1][0].size(0)
    dev = xyz.device
    out = torch.empty((B, cout, M), dtype=torch.float32, device=dev)
    out_pm = torch.empty((B, M, cout), dtype=torch.float32, device=dev) if want_pm else None
    idx_out = torch.empty((B, M, int(nsample)), dtype=torch.int32, device=dev) if (want_idx and idx is None) else None
    pp, pn = _plan_args(plan)
    with torch.cuda.device(dev):
        nbytes = int(_L().b200pn2_sa_forward_workspace(B, N, M, C, int(nsample), int(features_pm is not None),
                                                       int(idx is not None or idx_out is not None)))
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev) if nbytes else None
        cabi.check(_L().b200pn2_sa_forward_planned(B, N, M, C, float(radius), int(nsample), int(bool(use_xyz)),
                                                   int(bool(normalize_xyz)), _p(xyz), _p(features), _p(features_pm),
                                                   _p(new_xyz), _p(idx), len(layers), arr, _p(out), _p(out_pm),
                                                   _p(idx_out), _p(ws), nbytes, pp, pn, stream_ptr()), "sa_forward")
    return out, out_pm, (idx if idx is not None else idx_out)


def interp_mlp_forward(known_feats, idx3, weight3, rel_xyz, nsample, layers, known_feats_pm=None, plan=None):
    """Fused 3-neighbour blend -> cat(rel xyz) -> SharedMLP -> max over nsample (include/b200_pointnet2.h:
    b200pn2_interp_mlp_forward_planned).  known_feats (B,C,m), idx3/weight3/rel_xyz (B, M*nsample, 3) -> (B, cout, M)."""
    for t, nm in ((idx3, "idx3"), (weight3, "weight3")):
        _contig(t, nm); _cuda(t, nm)
    _is_int(idx3, "idx3"); _is_float(weight3, "weight3")
    if rel_xyz is not None:
        _contig(rel_xyz, "rel_xyz"); _is_float(rel_xyz, "rel_xyz"); _cuda(rel_xyz, "rel_xyz")
    if known_feats is not None:
        _contig(known_feats, "known_feats"); _is_float(known_feats, "known_feats"); _cuda(known_feats, "known_feats")
        B, C, m = known_feats.shape
    else:
        _contig(known_feats_pm, "known_feats_pm"); _is_float(known_feats_pm, "known_feats_pm")
        B, m, C = known_feats_pm.shape
    rows = idx3.size(1)
    _chk(rows % int(nsample) == 0, "idx3 rows must be a multiple of nsample")
    M = rows // int(nsample)
    arr, keep = _layer_array(layers)
    dev = idx3.device
    out = torch.empty((B, layers[-1][0].size(0), M), dtype=torch.float32, device=dev)
    nbytes = 0 if known_feats_pm is not None else ((B * m * C * 4 + 255) // 256) * 256
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev) if nbytes else None
    pp, pn = _plan_args(plan)
    with torch.cuda.device(dev):
        cabi.check(_L().b200pn2_interp_mlp_forward_planned(B, m, M, int(nsample), C, _p(known_feats), _p(known_feats_pm),
                                                           _p(idx3), _p(weight3), _p(rel_xyz), len(layers), arr, _p(out),
                                                           _p(ws), nbytes, pp, pn, stream_ptr()), "interp_mlp_forward")
    return out


def transpose_cn(x, ld=0):
    """(B,C,N) channel-major -> (B,N,ld or C) point-major, zero-padded columns (b200pn2_transpose_cn)."""
    _contig(x, "x"); _is_float(x, "x"); _cuda(x, None)
    B, C, N = x.shape
    width = int(ld) if ld else C
    out = torch.empty((B, N, width), dtype=torch.float32, device=x.device)
    if out.numel():
        with torch.cuda.device(x.device):
            cabi.check(_L().b200pn2_transpose_cn(B, C, N, _p(x), _p(out), width, stream_ptr()), "transpose_cn")
    return out


def fp_rows_forward(known_feats_pm, skip_feats_pm, idx3, weight3, layers, relu_last=True, want_cm=True, want_pm=False,
                    plan=None):
    """Feature-propagation rows [blend of 3 known rows | skip row] -> SharedMLP stack -> rows (b200pn2_fp_rows_forward).
    known_feats_pm (B,m,C2), skip_feats_pm (B,n,C1) or None, idx3/weight3 (B,n,3) -> (out (B,cout,n), out_pm (B,n,cout))."""
    for t, nm in ((known_feats_pm, "known_feats_pm"), (idx3, "idx3"), (weight3, "weight3")):
        _contig(t, nm); _cuda(t, nm)
    _is_float(known_feats_pm, "known_feats_pm"); _is_int(idx3, "idx3"); _is_float(weight3, "weight3")
    B, m, C2 = known_feats_pm.shape
    n = idx3.size(1)
    C1 = 0
    if skip_feats_pm is not None:
        _contig(skip_feats_pm, "skip_feats_pm"); _is_float(skip_feats_pm, "skip_feats_pm"); _cuda(skip_feats_pm, "skip_feats_pm")
        C1 = skip_feats_pm.size(2)
    arr, keep = _layer_array(layers)
    cout = layers[-1][0].size(0)
    dev = idx3.device
    out = torch.empty((B, cout, n), dtype=torch.float32, device=dev) if want_cm else None
    out_pm = torch.empty((B, n, cout), dtype=torch.float32, device=dev) if want_pm else None
    pp, pn = _plan_args(plan)
    with torch.cuda.device(dev):
        cabi.check(_L().b200pn2_fp_rows_forward(B, n, m, C2, C1, _p(known_feats_pm), _p(skip_feats_pm), _p(idx3),
                                                _p(weight3), len(layers), arr, int(bool(relu_last)), _p(out), _p(out_pm),
                                                pp, pn, stream_ptr()), "fp_rows_forward")
    return out, out_pm


def row_mlp_forward(x_pm, layers, relu_last=True, want_cm=True, want_pm=False, plan=None, channels=None):
    """1x1-conv stack on plain rows (b200pn2_row_mlp_forward).  x_pm (S,R,ld) point-major rows holding `channels` (default
    ld) channels -> (out (S,cout,R), out_pm (S,R,cout))."""
    _contig(x_pm, "x_pm"); _is_float(x_pm, "x_pm"); _cuda(x_pm, None)
    S, R, ld = x_pm.shape
    C = int(channels) if channels else ld
    arr, keep = _layer_array(layers)
    cout = layers[-1][0].size(0)
    dev = x_pm.device
    out = torch.empty((S, cout, R), dtype=torch.float32, device=dev) if want_cm else None
    out_pm = torch.empty((S, R, cout), dtype=torch.float32, device=dev) if want_pm else None
    pp, pn = _plan_args(plan)
    with torch.cuda.device(dev):
        cabi.check(_L().b200pn2_row_mlp_forward(S, R, C, ld, _p(x_pm), len(layers), arr, int(bool(relu_last)), _p(out),
                                                _p(out_pm), pp, pn, stream_ptr()), "row_mlp_forward")
    return out, out_pm


def row_mlp_forward_cm(x_cm, layers, relu_last=True, want_cm=True, want_pm=False, plan=None):
    """The same stack on channel-major rows (b200pn2_row_mlp_forward_cm): x_cm (S,C,R) -- a (B,C,H,W) conv input viewed
    (B,C,H*W) -- read in place, no transpose pass -> (out (S,cout,R), out_pm (S,R,cout))."""
    _contig(x_cm, "x_cm"); _is_float(x_cm, "x_cm"); _cuda(x_cm, None)
    S, C, R = x_cm.shape
    arr, keep = _layer_array(layers)
    cout = layers[-1][0].size(0)
    dev = x_cm.device
    out = torch.empty((S, cout, R), dtype=torch.float32, device=dev) if want_cm else None
    out_pm = torch.empty((S, R, cout), dtype=torch.float32, device=dev) if want_pm else None
    pp, pn = _plan_args(plan)
    with torch.cuda.device(dev):
        cabi.check(_L().b200pn2_row_mlp_forward_cm(S, R, C, _p(x_cm), len(layers), arr, int(bool(relu_last)), _p(out),
                                                   _p(out_pm), pp, pn, stream_ptr()), "row_mlp_forward_cm")
    return out, out_pm


def split_row_groups(layers):
    """Cut a stack into runs one tensor-core launch can take: every layer of a run but its last is a hidden layer
    (width a multiple of 32, <= 128), the last may be up to 256 wide; at most 4 layers per run."""
    groups, cur = [], []
    for w, sc, sh in layers:
        cur.append((w, sc, sh))
        cout = w.size(0)
        if not (cout % 32 == 0 and cout <= 128) or len(cur) == 4:
            groups.append(cur)
            cur = []
    if cur:
        groups.append(cur)
    return groups


# ---- training-mode fused SA (include/b200_pointnet2.h: b200pn2_sa_train_forward / _backward) -------------------------

def _bn_array(layers):
    """[(weight (cout,cin), gamma, beta, running_mean or None, running_var or None)] -> b200_bn_layer[]"""
    arr = (cabi.BnLayer * len(layers))()
    keep = []
    for i, (w, g, b, rm, rv) in enumerate(layers):
        w2 = w.reshape(w.size(0), -1)
        for t, nm in ((w2, "weight"), (g, "gamma"), (b, "beta")):
            _contig(t, nm); _is_float(t, nm); _cuda(t, nm)
        keep.append(w2)
        arr[i].cin, arr[i].cout = w2.size(1), w2.size(0)
        arr[i].weight, arr[i].gamma, arr[i].beta = w2.data_ptr(), g.data_ptr(), b.data_ptr()
        arr[i].running_mean = rm.data_ptr() if rm is not None else None
        arr[i].running_var = rv.data_ptr() if rv is not None else None
    return arr, keep


def sa_train_supported(C, use_xyz, layers, rows_mode=False):
    arr, keep = _bn_array(layers)
    return int(_L().b200pn2_sa_train_saved_bytes(1, 1, 1, int(C), int(bool(use_xyz)), len(layers), arr, int(rows_mode))) > 0


def sa_train_forward(xyz, features_pm, new_xyz, idx, radius, nsample, layers, eps, momentum, use_xyz=True,
                     normalize_xyz=False, x_rows=None, groups=None):
    """Training-mode SA MLP forward.  Gather mode: xyz (B,N,3), features_pm (B,N,C) or None, new_xyz (B,M,3), idx
    (B,M,nsample).  Rows mode: x_rows (G*nsample, C) with groups = (B, M).  layers as for _bn_array (running statistics
    are updated in place).  Returns (out (B,cout,M), saved uint8 tensor)."""
    arr, keep = _bn_array(layers)
    if x_rows is not None:
        _contig(x_rows, "x_rows"); _is_float(x_rows, "x_rows"); _cuda(x_rows, None)
        B, M = groups
        N, C, dev = 0, x_rows.size(1), x_rows.device
        _chk(x_rows.size(0) == B * M * int(nsample), "x_rows must hold B*M*nsample rows")
    else:
        for t, nm in ((xyz, "xyz"), (new_xyz, "new_xyz")):
            _contig(t, nm); _is_float(t, nm); _cuda(t, nm)
        _contig(idx, "idx"); _is_int(idx, "idx"); _cuda(idx, "idx")
        B, N, M, dev = xyz.size(0), xyz.size(1), new_xyz.size(1), xyz.device
        C = 0
        if features_pm is not None:
            _contig(features_pm, "features_pm"); _is_float(features_pm, "features_pm"); _cuda(features_pm, "features_pm")
            C = features_pm.size(2)
    rows_mode = int(x_rows is not None)
    cout = layers[-1][0].size(0)
    with torch.cuda.device(dev):
        nsaved = int(_L().b200pn2_sa_train_saved_bytes(B, M, int(nsample), C, int(bool(use_xyz)), len(layers), arr, rows_mode))
        _chk(nsaved > 0, "sa_train_forward: stack not supported by the fused training kernels")
        nws = int(_L().b200pn2_sa_train_workspace_bytes(B, M, int(nsample), C, int(bool(use_xyz)), len(layers), arr, rows_mode, 0))
        saved = torch.empty((nsaved,), dtype=torch.uint8, device=dev)
        ws = torch.empty((max(nws, 1),), dtype=torch.uint8, device=dev)
        out = torch.empty((B, cout, M), dtype=torch.float32, device=dev)
        cabi.check(_L().b200pn2_sa_train_forward(B, N, M, C, float(radius), int(nsample), int(bool(use_xyz)),
                                                 int(bool(normalize_xyz)), _p(xyz), _p(features_pm), _p(new_xyz), _p(idx),
                                                 _p(x_rows), len(layers), arr, float(eps), float(momentum), _p(out),
                                                 _p(saved), nsaved, _p(ws), nws, stream_ptr()), "sa_train_forward")
    return out, saved


def sa_train_backward(grad_out, saved, idx, nsample, layers, B, N, M, C, use_xyz=True, x_rows=None, want_input_grad=True):
    """Returns (grad_features (B,C,N) or grad_rows or None, [grad_weight], [grad_gamma], [grad_beta])."""
    arr, keep = _bn_array(layers)
    _contig(grad_out, "grad_out"); _is_float(grad_out, "grad_out"); _cuda(grad_out, None)
    dev = grad_out.device
    rows_mode = int(x_rows is not None)
    L = len(layers)
    gw = [torch.empty_like(l[0].reshape(l[0].size(0), -1)) for l in layers]
    gg = [torch.empty_like(l[1]) for l in layers]
    gb = [torch.empty_like(l[2]) for l in layers]
    pw = (ctypes.c_void_p * L)(*[t.data_ptr() for t in gw])
    pg = (ctypes.c_void_p * L)(*[t.data_ptr() for t in gg])
    pb = (ctypes.c_void_p * L)(*[t.data_ptr() for t in gb])
    gin = None
    if want_input_grad and C > 0:
        gin = torch.empty((x_rows.size(0), C) if rows_mode else (B, C, N), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        nws = int(_L().b200pn2_sa_train_workspace_bytes(B, M, int(nsample), C, int(bool(use_xyz)), L, arr, rows_mode, 1))
        ws = torch.empty((max(nws, 1),), dtype=torch.uint8, device=dev)
        cabi.check(_L().b200pn2_sa_train_backward(B, N, M, C, int(nsample), int(bool(use_xyz)), _p(idx), _p(x_rows), L, arr,
                                                  _p(grad_out), _p(saved), saved.numel(),
                                                  _p(gin) if (gin is not None and not rows_mode) else ctypes.c_void_p(0),
                                                  _p(gin) if (gin is not None and rows_mode) else ctypes.c_void_p(0),
                                                  pw, pg, pb, _p(ws), nws, stream_ptr()), "sa_train_backward")
    return gin, gw, gg, gb
