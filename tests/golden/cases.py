"""Input definitions shared by the golden-vector generator (make_golden.py, runs the reference's
own CUDA kernels on the GPU box) and the tests that consume tests/golden/*.npz.  Inputs are
regenerated from seeds; only outputs are stored."""
import numpy as np

from posecnn_b200 import synth

HOUGH_CASES = {
    # name: (scene kwargs, is_train, vote_thr, per_thr, skip)
    "h_test_a": (dict(batch=2, height=120, width=160, num_classes=5, objects_per_image=3, seed=3, min_pixels=520), 0, -1.0, 0.02, 10),
    "h_train_b": (dict(batch=2, height=120, width=160, num_classes=6, objects_per_image=3, seed=7, min_pixels=520), 1, -1.0, 0.02, 10),
    "h_thr_c": (dict(batch=2, height=120, width=160, num_classes=5, objects_per_image=3, seed=13, min_pixels=520, dir_noise=0.1), 0, 20.0, 0.002, 10),
    "h_skip1_d": (dict(batch=1, height=96, width=128, num_classes=4, objects_per_image=2, seed=17, min_pixels=520), 0, -1.0, 0.02, 1),
    "h_full_e": (dict(batch=1, height=480, width=640, num_classes=22, seed=1234), 0, -1.0, 0.02, 10),
}


def hough_inputs(name):
    kw, is_train, vt, pt, skip = HOUGH_CASES[name]
    return synth.make_scene(**kw), is_train, vt, pt, skip


def roi_inputs():
    rng = np.random.default_rng(0)
    data = rng.standard_normal((3, 30, 40, 32)).astype(np.float32)
    rois = synth.make_rois(17, 3, num_classes=22, seed=2)
    rois[3, 2:6] = [700, 500, 710, 505]
    rois[4, 2:6] = [100, 100, 90, 80]
    grad = rng.standard_normal((17, 7, 7, 32)).astype(np.float32)
    return data, rois, grad


def hard_label_inputs():
    rng = np.random.default_rng(2)
    return rng.random((2, 37, 53, 22)).astype(np.float32), rng.integers(-1, 22, (2, 37, 53)).astype(np.int32)


def projection_inputs():
    case = synth.make_projection_case(2, 48, 64, 8, 3, 16, seed=5)
    rng = np.random.default_rng(6)
    case["g3"] = rng.standard_normal((2, 16, 16, 16, 8)).astype(np.float32)
    case["g2"] = rng.standard_normal((2, 48, 64, 8)).astype(np.float32)
    return case


def avgdist_inputs():
    pts = synth.make_model_points(22, 300)
    pred, targ, wt = synth.make_pose_batch(9, 22, seed=9)
    pred[0] = 0; targ[0] = 0; wt[0] = 0
    q = np.array([0.8, 0.2, -0.4, 0.4], np.float32); p = q + np.array([0.1, -0.05, 0.02, 0.07], np.float32)
    targ[0, 64:68] = q / np.linalg.norm(q); pred[0, 64:68] = p / np.linalg.norm(p); wt[0, 64:68] = 1
    return pred, targ, wt, pts, synth.LOV_SYMMETRY.copy()


def nms_inputs(seed=77, n=96, num_classes=22, batch=1):
    """Hough-style ROI rows [batch, cls, x1, y1, x2, y2, score] with many same-class overlaps and DISTINCT scores (the
    reference's argsort tie order is numpy-version dependent), + poses_init [n,7] and poses_pred [n,4C]."""
    rng = np.random.default_rng(seed)
    cls = rng.integers(1, 6, size=n)                      # few classes -> same-class overlaps are common
    cx = rng.uniform(100, 540, size=n); cy = rng.uniform(80, 400, size=n)
    # cluster boxes around a few centres per class so that IoU > 0.5 happens often
    anchor = rng.integers(0, 12, size=n)
    acx = rng.uniform(120, 520, size=12); acy = rng.uniform(100, 380, size=12)
    near = rng.random(n) < 0.7
    cx = np.where(near, acx[anchor] + rng.normal(0, 12, n), cx); cy = np.where(near, acy[anchor] + rng.normal(0, 12, n), cy)
    w = rng.uniform(60, 160, size=n); h = rng.uniform(60, 160, size=n)
    rois = np.zeros((n, 7), np.float32)
    rois[:, 0] = rng.integers(0, batch, size=n)
    rois[:, 1] = cls
    rois[:, 2] = cx - w / 2; rois[:, 3] = cy - h / 2; rois[:, 4] = cx + w / 2; rois[:, 5] = cy + h / 2
    rois[:, 6] = rng.permutation(n).astype(np.float32) * 7 + 500       # distinct integer-valued vote counts
    poses_init = rng.standard_normal((n, 7)).astype(np.float32)
    poses_pred = np.tanh(rng.standard_normal((n, 4 * num_classes))).astype(np.float32)
    return rois, poses_init, poses_pred


def vertex_target_inputs(seed=41, H=60, W=80, C=6, B=2):
    """label maps with a few elliptical objects, per-class projected centres and depths; one class present in the label
    map but missing from cls_indexes (no targets), one listed class without pixels."""
    rng = np.random.default_rng(seed)
    label = np.zeros((B, H, W), np.int32)
    centers = np.zeros((B, C, 3), np.float32)            # (cx, cy, z); z = 0 -> class not in cls_indexes
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        classes = rng.permutation(np.arange(1, C))[:3]
        for c in classes:
            cx, cy = rng.uniform(0.2 * W, 0.8 * W), rng.uniform(0.2 * H, 0.8 * H)
            a, bb = rng.uniform(6, 14), rng.uniform(6, 14)
            label[b][((xx - cx) / a) ** 2 + ((yy - cy) / bb) ** 2 <= 1] = c
            centers[b, c] = (cx + rng.normal(0, 2), cy + rng.normal(0, 2), rng.uniform(0.5, 1.5))
        centers[b, classes[0], 2] = 0.0                  # labelled but not listed
        spare = [c for c in range(1, C) if c not in classes]
        if spare:
            centers[b, spare[0]] = (10.0, 10.0, 0.9)     # listed but without pixels
        if b == 0:
            centers[0, classes[1], 0:2] = (17.0, 23.0)   # integer centre on a pixel: zero vector / (0 + 1e-10)
            label[0, 23, 17] = classes[1]
    return label, centers


def vertex_target_multi_inputs(seed=43, H=60, W=80, C=6, B=2, I=5):
    """Multi-instance images (minibatch.py:425-431): two instances of one class told apart by an instance-mask image.
    Returns label [B,H,W] int32, mask [B,H,W] int32 and instances [B,I,5] f32 = (cls, mask id, cx, cy, z); z = 0 marks an
    unused slot.  Image 0 also has an instance whose mask region carries a DIFFERENT label (no pixels: skipped) and a
    repeated (cls, id) pair (the later instance overwrites the earlier one)."""
    rng = np.random.default_rng(seed)
    label = np.zeros((B, H, W), np.int32)
    mask = np.zeros((B, H, W), np.int32)
    inst = np.zeros((B, I, 5), np.float32)
    yy, xx = np.mgrid[0:H, 0:W]
    for b in range(B):
        classes = [2, 2, 4, 1][: I - 1]                         # class 2 twice
        for i, c in enumerate(classes):
            cx, cy = rng.uniform(0.15 * W, 0.85 * W), rng.uniform(0.15 * H, 0.85 * H)
            a, bb = rng.uniform(6, 12), rng.uniform(6, 12)
            region = ((xx - cx) / a) ** 2 + ((yy - cy) / bb) ** 2 <= 1
            label[b][region] = c
            mask[b][region] = i + 1                               # cls_indexes_old[i] + 1
            inst[b, i] = (c, i + 1, cx + rng.normal(0, 2), cy + rng.normal(0, 2), rng.uniform(0.5, 1.5))
    label[0][mask[0] == 4] = 3                                    # instance 3 (class 1): its mask region is labelled 3 -> no pixels
    inst[0, 4] = inst[0, 0]; inst[0, 4, 2:5] = (33.0, 21.0, 1.25)  # same (cls, id) as instance 0, listed later: it wins
    return label, mask, inst
