"""Seeded synthetic inputs shared by bench.py, the tests (tests/cases.py re-exports this module) and the golden-vector
generator.

Shapes follow SURVEY.md section 8(d): C1 = one 2000-point cloud for PointnetSAModuleVotes(npoint=128, r=0.2, ns=32,
mlp=[4,32]); C2 = ScanNet-shaped (B, 40000, 4) scenes; C3 = 256 x 256 rotated boxes.  Adversarial variants cover the
edge cases of the reference kernels: duplicate points (exact distance ties -> FPS bit-reversed tie order), points
inside the FPS origin-skip sphere (x^2+y^2+z^2 <= 1e-3), empty balls, more neighbours than nsample, ragged sizes.
"""
import numpy as np


def cloud(seed, B, N, extent=(4.0, 4.0, 2.5), dup_frac=0.0, origin_frac=0.0, centre=True):
    """(B,N,3) float32 room-sized cloud; optional duplicated points and points near the origin."""
    rng = np.random.default_rng(seed)
    xyz = rng.random((B, N, 3), dtype=np.float32) * np.asarray(extent, np.float32)
    if centre:
        xyz -= np.asarray(extent, np.float32) / 2  # origin inside the room, so the skip sphere matters
    nd = int(N * dup_frac)
    no = int(N * origin_frac)
    for b in range(B):
        if nd:
            src = rng.integers(0, N, nd)
            dst = rng.integers(0, N, nd)
            xyz[b, dst] = xyz[b, src]
        if no:
            where = rng.integers(0, N, no)
            xyz[b, where] = (rng.random((no, 3), dtype=np.float32) - 0.5) * 0.03  # |p|^2 < 1e-3 for most
    return np.ascontiguousarray(xyz, np.float32)


def scene_cloud(seed, B, N, room=(8.0, 8.0, 3.0)):
    """ScanNet-like scenes: 70% of the points on 6 random axis-aligned planes, 30% volumetric;
    returns (B,N,4) = xyz + height above the 1st-percentile floor."""
    rng = np.random.default_rng(seed)
    pc = np.empty((B, N, 4), np.float32)
    room = np.asarray(room, np.float32)
    for b in range(B):
        xyz = rng.random((N, 3), dtype=np.float32) * room
        n_surf = int(0.7 * N)
        plane = rng.integers(0, 6, n_surf)
        for p in range(6):
            sel = np.nonzero(plane == p)[0]
            axis = p % 3
            xyz[sel, axis] = np.float32(rng.random() * room[axis])
        xyz = xyz[rng.permutation(N)]
        xyz -= np.array([room[0] / 2, room[1] / 2, 0], np.float32)
        floor = np.percentile(xyz[:, 2], 1)
        pc[b, :, :3] = xyz
        pc[b, :, 3] = xyz[:, 2] - floor
    return pc


def boxes(seed, n, extent=(8.0, 8.0, 3.0), jitter_of=None, jitter=0.1):
    """(n,7) [x,y,z,dx,dy,dz,heading]; with `jitter_of` the boxes are noisy copies (non-trivial IoUs)."""
    rng = np.random.default_rng(seed)
    if jitter_of is not None:
        out = jitter_of + rng.normal(0, jitter, jitter_of.shape).astype(np.float32)
        out[:, 3:6] = np.abs(out[:, 3:6]) + 0.05
        return np.ascontiguousarray(out, np.float32)
    c = rng.random((n, 3), dtype=np.float32) * np.asarray(extent, np.float32)
    s = rng.random((n, 3), dtype=np.float32) * 1.5 + 0.2
    h = (rng.random((n, 1), dtype=np.float32) - 0.5) * 2 * np.pi
    return np.ascontiguousarray(np.concatenate([c, s, h], 1), np.float32)


def aabb_boxes(seed, K, ncls, extent=(4.0, 4.0, 2.0)):
    """(K,8) float64 rows [x1,y1,z1,x2,y2,z2,score,class] with float32-representable values and DISTINCT scores
    (the layout of the reference's boxes_3d_with_prob, models/ap_helper.py:187-197)."""
    rng = np.random.default_rng(1000 + seed)
    c = rng.random((K, 3)) * np.asarray(extent)
    s = rng.random((K, 3)) * 1.5 + 0.05
    b = np.zeros((K, 8))
    b[:, 0:3] = (c - s / 2).astype(np.float32)
    b[:, 3:6] = (c + s / 2).astype(np.float32)
    b[:, 6] = ((rng.permutation(K) + rng.random(K) * 0.5) / K).astype(np.float32)
    b[:, 7] = rng.integers(0, ncls, K)
    return b


def degenerate_boxes():
    """Identical, touching, nested, zero-size, axis multiples, near-coincident rectangles."""
    b = [
        [0, 0, 0, 2, 2, 1, 0],                       # 0 reference square
        [0, 0, 0, 2, 2, 1, 0],                       # 1 identical
        [1, 0, 0, 2, 2, 1, 0],                       # 2 shifted by half
        [2, 0, 0, 2, 2, 1, 0],                       # 3 touching edge
        [2.005, 0, 0, 2, 2, 1, 0],                   # 4 separated by less than the 1e-2 corner margin
        [5, 5, 0, 2, 2, 1, 0],                       # 5 disjoint
        [0, 0, 0, 2, 2, 1, np.pi / 4],               # 6 45 degrees
        [0, 0, 0, 2, 2, 1, np.pi / 2],               # 7 90 degrees
        [0, 0, 0, 2, 2, 1, np.pi],                   # 8 180 degrees
        [0, 0, 0, 0.5, 0.5, 1, 0.3],                 # 9 nested small
        [0, 0, 0, 0, 0, 0, 0],                       # 10 zero size
        [0, 0, 0.9, 2, 2, 1, 0],                     # 11 height overlap 0.1
        [0, 0, 2, 2, 2, 1, 0],                       # 12 no height overlap
        [0.3, -0.2, 0.1, 3, 1, 2, 1.1],              # 13 generic
        [0.3, -0.2, 0.1, 3, 1, 2, 1.1000001],        # 14 near-coincident
        [-1000, -1000, -1000, 1, 1, 1, 0],           # 15 padded GT slot (loss_helper_iou.py:56-58)
        [0, 0, 0, 4, 0.2, 1, 0.7],                   # 16 thin
        [0, 0, 0, 0.2, 4, 1, -0.7],                  # 17 thin crossing
    ]
    return np.asarray(b, np.float32)


def mlp_params(seed, spec, bn=True):
    """Random SharedMLP parameters in eval-BN form: list of dicts (weight, gamma, beta, mean, var)."""
    rng = np.random.default_rng(seed)
    layers = []
    for cin, cout in zip(spec[:-1], spec[1:]):
        w = (rng.standard_normal((cout, cin)) * np.sqrt(2.0 / cin)).astype(np.float32)
        layers.append(dict(weight=w,
                           gamma=(rng.random(cout) * 0.5 + 0.75).astype(np.float32),
                           beta=(rng.standard_normal(cout) * 0.1).astype(np.float32),
                           mean=(rng.standard_normal(cout) * 0.1).astype(np.float32),
                           var=(rng.random(cout) * 0.5 + 0.5).astype(np.float32)))
    return layers


def scene_labels(seed, pc, num_class, num_heading_bin, num_size_cluster, mean_size_arr, max_gt=64, extent=(6.0, 6.0, 2.0)):
    """Ground-truth tensors in the reference's dataset format (scannet/scannet_detection_dataset.py:66-190,
    sunrgbd/sunrgbd_detection_dataset.py:78-220) for synthetic scenes `pc` (B,N,3+C): 3-12 boxes per scene in 64 padded
    slots, size = mean_size[class] + residual, heading = class * 2pi/NH + residual, votes = offset to the
    centre of the (first) axis-aligned-containing box, repeated three times.  Returns a dict of numpy arrays."""
    rng = np.random.default_rng(seed + 100)
    B, N = pc.shape[0], pc.shape[1]
    mean_size_arr = np.asarray(mean_size_arr, np.float32)
    out = dict(center_label=np.zeros((B, max_gt, 3), np.float32),
               heading_class_label=np.zeros((B, max_gt), np.int64),
               heading_residual_label=np.zeros((B, max_gt), np.float32),
               size_class_label=np.zeros((B, max_gt), np.int64),
               size_residual_label=np.zeros((B, max_gt, 3), np.float32),
               sem_cls_label=np.zeros((B, max_gt), np.int64),
               box_label_mask=np.zeros((B, max_gt), np.float32),
               vote_label=np.zeros((B, N, 9), np.float32),
               vote_label_mask=np.zeros((B, N), np.int64))
    per = 2 * np.pi / num_heading_bin
    for b in range(B):
        n = int(rng.integers(3, 13))
        c = rng.random((n, 3), dtype=np.float32) * np.asarray(extent, np.float32)
        c[:, 0:2] -= np.asarray(extent[0:2], np.float32) / 2
        cls = rng.integers(0, num_size_cluster, n)
        res = (rng.standard_normal((n, 3)) * 0.1).astype(np.float32) * mean_size_arr[cls]
        out["center_label"][b, :n] = c
        out["size_class_label"][b, :n] = cls
        out["size_residual_label"][b, :n] = res
        out["sem_cls_label"][b, :n] = cls % num_class
        out["box_label_mask"][b, :n] = 1.0
        if num_heading_bin > 1:
            hc = rng.integers(0, num_heading_bin, n)
            out["heading_class_label"][b, :n] = hc
            out["heading_residual_label"][b, :n] = ((rng.random(n) - 0.5) * per * 0.9).astype(np.float32)
        half = (mean_size_arr[cls] + res) / 2
        xyz = pc[b, :, :3]
        for i in range(n - 1, -1, -1):  # the first containing box wins
            inside = np.all(np.abs(xyz - c[i]) <= half[i], axis=1)
            out["vote_label"][b, inside] = np.tile(c[i] - xyz[inside], (1, 3))
            out["vote_label_mask"][b, inside] = 1
    return out
