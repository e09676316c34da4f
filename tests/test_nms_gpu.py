"""GPU parity of the device-side NMS + pose assembly (C ABI pcnn_nms_pose_fwd) against the reference's own
lib/utils/nms.py outputs (tests/golden/nms.npz) and the numpy oracle (ties, batches, row counts)."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle
from tests.golden import cases
from tests.util import to_np

pytestmark = pytest.mark.gpu
T = lambda a, dev: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("tag,kw", [("a", dict(seed=77, n=96)), ("b", dict(seed=78, n=128)), ("c", dict(seed=79, n=7)),
                                    ("d", dict(seed=80, n=1))])
def test_nms_pose_matches_reference_golden(cuda, tag, kw):
    from posecnn_b200.utils import nms as dev_nms
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nms.npz"))
    rois, poses_init, poses_pred = cases.nms_inputs(**kw)
    keep, r, p, nk = dev_nms.nms_pose_capacity(T(rois, cuda), T(poses_init, cuda), T(poses_pred, cuda), None, 0.5, per_image=False)
    n = int(nk.item())
    np.testing.assert_array_equal(to_np(keep)[:n], g[f"{tag}_keep"])          # index work: exact
    np.testing.assert_array_equal(to_np(r)[:n], g[f"{tag}_rois"])
    np.testing.assert_array_equal(to_np(p)[:n], g[f"{tag}_poses"])
    assert (to_np(keep)[n:] == -1).all() and not to_np(r)[n:].any() and not to_np(p)[n:].any()
    np.testing.assert_array_equal(to_np(dev_nms.nms(T(rois, cuda), 0.5)), g[f"{tag}_keep"])   # lib/utils/nms.py signature


def test_nms_ties_batches_and_row_count(cuda):
    """Vote-count scores tie often; canonical order = stable argsort reversed.  per_image keeps images independent.
    num_rois on the device limits the rows like Hough's count (0 -> the dummy row)."""
    from posecnn_b200.utils import nms as dev_nms
    rois, poses_init, poses_pred = cases.nms_inputs(seed=5, n=300, batch=4)
    rois[:, 6] = np.floor(rois[:, 6] / 50)                # heavy ties
    for per_image in (False, True):
        want_keep = oracle.nms(rois, 0.5, per_image=per_image)
        wr, wp = oracle.assemble_poses(rois, poses_init, poses_pred, want_keep)
        keep, r, p, nk = dev_nms.nms_pose_capacity(T(rois, cuda), T(poses_init, cuda), T(poses_pred, cuda), None, 0.5, per_image=per_image)
        n = int(nk.item())
        assert n == len(want_keep)
        np.testing.assert_array_equal(to_np(keep)[:n], np.asarray(want_keep, np.int32))
        np.testing.assert_array_equal(to_np(r)[:n], wr)
        np.testing.assert_array_equal(to_np(p)[:n], wp)
    assert len(oracle.nms(rois, 0.5, per_image=True)) > len(oracle.nms(rois, 0.5, per_image=False))
    for count in (0, 1, 17, 300):
        nr = torch.tensor([count], dtype=torch.int32, device=cuda)
        m = max(count, 1)
        want_keep = oracle.nms(rois[:m], 0.5, per_image=True)
        keep, r, p, nk = dev_nms.nms_pose_capacity(T(rois, cuda), T(poses_init, cuda), None, nr, 0.5, per_image=True)
        n = int(nk.item())
        np.testing.assert_array_equal(to_np(keep)[:n], np.asarray(want_keep, np.int32))
        np.testing.assert_array_equal(to_np(p)[:n], poses_init[want_keep])     # no poses_pred: poses_init passes through
    # capacity limit and argument checks are errors, not crashes
    big = torch.zeros((1153, 7), device=cuda)
    with pytest.raises(RuntimeError):
        dev_nms.nms_pose_capacity(big, big, None, None, 0.5)


def test_network_detections(cuda):
    """vgg16_convs.forward publishes the post-NMS records; they equal oracle NMS + assembly on the network's own
    rois / poses_init / poses_tanh."""
    from posecnn_b200 import synth
    from posecnn_b200.networks.vgg16_convs import vgg16_convs
    net = vgg16_convs(num_classes=6, device=cuda, vote_threshold=5.0).init_random(seed=0, bias_std=0.05)   # threshold mode: several maxima per class
    rgb, _ = synth.make_images(2, 64, 96, seed=3)
    data = torch.from_numpy(rgb).to(cuda)
    meta = torch.from_numpy(np.stack([synth.make_meta(synth.intrinsics(64, 96))] * 2)).to(cuda)
    ext = torch.from_numpy(synth.extents_for(6)).to(cuda)
    out = net.forward(data, meta, ext)
    rois, pi, pt = to_np(out["rois"]), to_np(out["poses_init"]), to_np(out["poses_tanh"])
    keep = oracle.nms(rois, 0.5, per_image=True)
    wr, wp = oracle.assemble_poses(rois, pi, pt, keep)
    n = int(out["num_detections"].item())
    assert n == len(keep)
    np.testing.assert_array_equal(to_np(out["detections_rois"])[:n], wr)
    np.testing.assert_array_equal(to_np(out["detections_poses"])[:n], wp)
