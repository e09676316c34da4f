"""Gradient wiring of RoiPool — mirrors lib/roi_pooling_layer/roi_pooling_op_grad.py:29-50."""
import torch

from . import roi_pooling_op


class RoiPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, rois, pooled_height, pooled_width, spatial_scale, pool_channel):
        top, argmax = roi_pooling_op.roi_pool(data, rois, pooled_height, pooled_width, spatial_scale, pool_channel)
        ctx.save_for_backward(data, rois, argmax)
        ctx.attrs = (pooled_height, pooled_width, spatial_scale, pool_channel)
        ctx.mark_non_differentiable(argmax)
        return top, argmax

    @staticmethod
    def backward(ctx, grad, _):
        data, rois, argmax = ctx.saved_tensors
        g = roi_pooling_op.roi_pool_grad(data, rois, argmax, grad.contiguous(), *ctx.attrs)
        return g, None, None, None, None, None


def roi_pool(data, rois, pooled_height, pooled_width, spatial_scale, pool_channel=0):
    return RoiPool.apply(data, rois, pooled_height, pooled_width, spatial_scale, pool_channel)
