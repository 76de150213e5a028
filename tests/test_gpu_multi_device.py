"""GPU (needs two devices; skipped otherwise): one process, one python thread per device -- the way the reference's
nn.DataParallel replicas call the operators (train.py:187-191).  Every host-side cache of the library that the CUDA
runtime keeps per device (the > 48 KB dynamic shared-memory opt-ins, SM counts, the NMS workspace, scratch pools) must be
per device: the second device's first launch of each kernel may not inherit the first device's state."""
import importlib
import threading

import numpy as np
import pytest
import torch

import cases

pytestmark = pytest.mark.gpu


def _work(dev_index, xyz, feats, layers, boxes, scores, out, errors):
    try:
        import pointnet2._ext as ext
        from pcdet.ops.iou3d_nms import iou3d_nms_utils as iu
        dev = torch.device("cuda", dev_index)
        with torch.cuda.device(dev):
            x = torch.from_numpy(xyz).to(dev)
            f = torch.from_numpy(feats).to(dev)
            trip = [(torch.from_numpy(w).to(dev), torch.from_numpy(s).to(dev), torch.from_numpy(h).to(dev)) for w, s, h in layers]
            res = []
            for _ in range(3):
                inds = ext.furthest_point_sampling(x, 512)                                      # cluster FPS (non-default smem size)
                new_xyz = ext.gather_points(x.transpose(1, 2).contiguous(), inds).transpose(1, 2).contiguous()
                o, _, idx = ext.sa_forward(x, f, new_xyz, 0.4, 32, trip, normalize_xyz=True, want_idx=True)   # tcgen05 kernel, ~200 KB smem
                iou = iu.boxes_iou3d_gpu(torch.from_numpy(boxes).to(dev), torch.from_numpy(boxes[::-1].copy()).to(dev))
                keep, _ = iu.nms_gpu(torch.from_numpy(boxes).to(dev), torch.from_numpy(scores).to(dev), 0.25)
                res = [inds.cpu(), idx.cpu(), o.cpu(), iou.cpu(), keep.cpu()]
            torch.cuda.synchronize(dev)
            out[dev_index] = res
    except Exception as e:  # noqa: BLE001
        errors.append((dev_index, repr(e)))


def test_two_devices_two_threads(pkg, orc):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import torch_ref
    xyz = cases.scene_cloud(5, 2, 9000)[:, :, :3].copy()
    feats = np.random.default_rng(6).standard_normal((2, 64, 9000)).astype(np.float32)
    layers = torch_ref.fold(cases.mlp_params(7, [67, 64, 64, 128]))
    boxes = cases.boxes(0, 300)
    scores = np.random.default_rng(8).random(300).astype(np.float32)
    out, errors = {}, []
    # device 1 FIRST on its own, so that no cache can have been primed by device 0; then both concurrently, twice
    _work(1, xyz, feats, layers, boxes, scores, out, errors)
    assert not errors, errors
    first = out[1]
    for _ in range(2):
        threads = [threading.Thread(target=_work, args=(d, xyz, feats, layers, boxes, scores, out, errors)) for d in (0, 1)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        for a, b, c in zip(out[0], out[1], first):
            assert torch.equal(a, b) and torch.equal(a, c)
    assert np.array_equal(out[0][0].numpy(), orc.furthest_point_sampling(xyz, 512))
