"""Gradient wiring of Houghvotinggpu — mirrors lib/hough_voting_gpu_layer/hough_voting_gpu_op_grad.py:18-35
(RegisterGradient: zeros for label and vertex, None for extents / meta_data / gt)."""
import torch

from . import hough_voting_gpu_op


class HoughVotingGPU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, label, vertex, extents, meta_data, gt, is_train, threshold_vote, threshold_percentage, skip_pixels):
        ctx.save_for_backward(label, vertex)
        outs = hough_voting_gpu_op.hough_voting_gpu(label, vertex, extents, meta_data, gt, is_train, threshold_vote,
                                                    threshold_percentage, skip_pixels)
        ctx.mark_non_differentiable(outs[4])
        return outs

    @staticmethod
    def backward(ctx, grad, tmp, tmp1, tmp2, _):
        label, vertex = ctx.saved_tensors
        _, g_vertex = hough_voting_gpu_op.hough_voting_gpu_grad(label, vertex, grad)
        # label is int32 (not differentiable in torch); the reference returns float zeros for it
        return None, g_vertex, None, None, None, None, None, None, None


def hough_voting_gpu(label, vertex, extents, meta_data, gt, is_train, threshold_vote, threshold_percentage, skip_pixels):
    return HoughVotingGPU.apply(label, vertex, extents, meta_data, gt, is_train, threshold_vote, threshold_percentage,
                                skip_pixels)
