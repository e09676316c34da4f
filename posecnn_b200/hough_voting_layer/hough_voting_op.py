"""Houghvoting (CPU RANSAC op) — import-compatible placeholder for lib/hough_voting_layer/hough_voting_op.py.

`lib/networks/network.py:17,253-254` imports this module unconditionally, so it has to exist for the reference's
network code to import against this package.  The op itself is the reference's CPU path: it is
restated only as baseline / test infrastructure (`oracle/cpu_hough_ransac.cpp`) and is deliberately NOT part of the
product — there is no CPU fallback.  Calling it fails loudly and points at the GPU op.
"""
from __future__ import annotations


def hough_voting(bottom_label, bottom_vertex, bottom_extents, bottom_meta_data, bottom_gt, is_train, name=None):
    raise NotImplementedError("Houghvoting (CPU RANSAC, lib/hough_voting_layer) is baseline-only in posecnn_b200; use "
                              "hough_voting_gpu_layer.hough_voting_gpu_op.hough_voting_gpu (Houghvotinggpu) instead")


def hough_voting_grad(bottom_label, bottom_vertex, grad, name=None):
    raise NotImplementedError("HoughvotingGrad is baseline-only in posecnn_b200; see hough_voting_gpu_layer")
