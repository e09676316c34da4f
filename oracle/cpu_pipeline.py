"""The reference's hot path on host cores, end to end — the CPU arm of `bench.py --impl reference`.

TEST INFRASTRUCTURE / CPU BASELINE ONLY: never imported by the product package.

The reference itself cannot run here (TensorFlow 1.x, Python 2, OpenCV C++), so the arm is the
oracle restatement of the SAME path the GPU arm times, executed in the reference's own op order:

    BGR - PIXEL_MEANS                                   lib/fcn/test.py:37-110
    VGG16 trunk, FCN heads (dense conv2d_transpose)     oracle/ref_network.py  (lib/networks/vgg16_convs.py:79-163; torch fp32 = the
                                                        TF/cuDNN arithmetic restated, all host threads)
    hough_voting (CPU RANSAC op)                        oracle/cpu_hough_ransac.cpp  (lib/hough_voting_layer/hough_voting_op.cc:88-857)
    roi_pool x2 + add, fc6-fc8, tanh                    oracle/posecnn_oracle.c + torch fp32  (vgg16_convs.py:177-197)
    nms + pose assembly                                 oracle.nms / assemble_poses  (lib/utils/nms.py:3-32, lib/fcn/test.py:197-211)

Same synthetic images, same seeded Kaiming weights and the same background calibration as the GPU arm.
"""
from __future__ import annotations

import math
import os
import time

import numpy as np
import torch

from . import cpu_hough, oracle, ref_network as R

PIXEL_MEANS = (102.9801, 115.9465, 122.7717)  # lib/fcn/config.py:242


def param_shapes(C=22, U=64):
    """Parameter table of vgg16_convs (COLOR): names and TF shapes in creation order (lib/networks/vgg16_convs.py:80-197)."""
    shapes = {}
    for item in R.VGG_CFG:
        if isinstance(item, tuple):
            name, ci, co = item
            shapes[f"{name}/weights"] = (3, 3, ci, co)
            shapes[f"{name}/biases"] = (co,)
    for name, co in (("score_conv5", U), ("score_conv4", U), ("score_conv5_vertex", 128), ("score_conv4_vertex", 128)):
        shapes[f"{name}/weights"] = (1, 1, 512, co)
        shapes[f"{name}/biases"] = (co,)
    shapes["score/weights"] = (1, 1, U, C); shapes["score/biases"] = (C,)
    shapes["vertex_pred/weights"] = (1, 1, 128, 3 * C); shapes["vertex_pred/biases"] = (3 * C,)
    shapes["fc6/weights"] = (7 * 7 * 512, 4096); shapes["fc6/biases"] = (4096,)
    shapes["fc7/weights"] = (4096, 4096); shapes["fc7/biases"] = (4096,)
    shapes["fc8/weights"] = (4096, 4 * C); shapes["fc8/biases"] = (4 * C,)
    return shapes


def init_random(C=22, seed=0):
    """The GPU arm's initialisation (seeded Kaiming-normal, zero biases) drawn from the same CPU generator stream."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    params = {}
    for name, shp in param_shapes(C).items():
        if name.endswith("weights"):
            params[name] = torch.randn(shp, generator=g) * math.sqrt(2.0 / int(np.prod(shp[:-1])))
        else:
            params[name] = torch.zeros(shp)
    return params


def forward(params, images_u8, meta, extents, C=22, threads=1, timings=None):
    """images_u8 [B,H,W,3] uint8 BGR -> (rois [N,7], poses [N,7]) after NMS, like lib/fcn/test.py:190-211."""
    t = [time.perf_counter()]
    x = (torch.from_numpy(images_u8).float() - torch.tensor(PIXEL_MEANS)).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        feats = R.trunk(params, x)
        t.append(time.perf_counter())
        score, label, prob, vertex = R.heads(params, feats["conv4_3"], feats["conv5_3"], C)
    t.append(time.perf_counter())
    lab = label.numpy().astype(np.int32)
    ver = vertex.permute(0, 2, 3, 1).contiguous().numpy().copy()
    # the CPU op reads the third vertex channel as a metric distance (no exp, ransac.h:105-116) and rejects negative
    # values (hough_voting_op.cc:338-352): give it exp(log z), its own convention, so RANSAC is timed, not the rejection loop
    ver[..., 2::3] = np.exp(np.clip(ver[..., 2::3], -10.0, 10.0))
    box, pose = cpu_hough.hough_voting(lab, ver, extents, meta, is_train=0, threads=threads)
    t.append(time.perf_counter())
    valid = box[:, 1] >= 0
    rois = np.concatenate([box, np.ones((box.shape[0], 1), np.float32)], 1)[valid]   # score column for nms
    poses_init = pose[valid]
    if rois.shape[0] == 0:
        if timings is not None:
            timings.append(np.diff(t + [time.perf_counter()]))
        return rois, poses_init
    c5 = feats["conv5_3"].permute(0, 2, 3, 1).contiguous().numpy()
    c4 = feats["conv4_3"].permute(0, 2, 3, 1).contiguous().numpy()
    p5, _ = oracle.roi_pool(c5, rois, 7, 7, 1.0 / 16.0)
    p4, _ = oracle.roi_pool(c4, rois, 7, 7, 1.0 / 8.0)
    with torch.no_grad():
        h = torch.from_numpy(p5 + p4).reshape(rois.shape[0], -1)
        h = torch.relu(h @ params["fc6/weights"] + params["fc6/biases"])
        h = torch.relu(h @ params["fc7/weights"] + params["fc7/biases"])
        poses_tanh = torch.tanh(h @ params["fc8/weights"] + params["fc8/biases"]).numpy()
    keep = oracle.nms(rois, 0.5, per_image=True)
    out_rois, out_poses = oracle.assemble_poses(rois, poses_init, poses_tanh, keep)
    t.append(time.perf_counter())
    if timings is not None:
        timings.append(np.diff(t))
    return out_rois, out_poses


def calibrate_background(params, images_u8, C=22, background_fraction=0.75):
    """Same harness rule as the GPU arm (vgg16_convs.calibrate_background): shift score/biases[0] so that about
    `background_fraction` of the pixels are labelled background.  Closed form on the probe frames: pixel p is background
    iff relu(s0 + shift) >= max_{c>0} relu(s_c) (arg-max ties go to the lowest index)."""
    x = (torch.from_numpy(images_u8).float() - torch.tensor(PIXEL_MEANS)).permute(0, 3, 1, 2).contiguous()
    with torch.no_grad():
        feats = R.trunk(params, x)
        s5 = R.conv(feats["conv5_3"], params["score_conv5/weights"], params["score_conv5/biases"])
        s4 = R.conv(feats["conv4_3"], params["score_conv4/weights"], params["score_conv4/biases"])
        up = R.deconv(s4 + R.deconv(s5, 4, 2), 16, 8)
        w = params["score/weights"].permute(3, 2, 0, 1)
        pre = torch.nn.functional.conv2d(up, w, params["score/biases"])
    rest = torch.relu(pre[:, 1:]).max(dim=1).values
    need = (rest - pre[:, 0]).flatten()            # smallest shift that makes the pixel background
    shift = float(torch.quantile(need[:: max(1, need.numel() // 1_000_000)], background_fraction))
    params["score/biases"][0] += shift
    return shift


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
