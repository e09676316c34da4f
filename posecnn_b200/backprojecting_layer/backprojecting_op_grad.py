"""Gradient wiring of Backproject — mirrors lib/backprojecting_layer/backprojecting_op_grad.py
(gradient for bottom_data only)."""
import torch

from . import backprojecting_op


class Backproject(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, label, depth, meta, label_3d, grid_size, kernel_size, threshold):
        ctx.save_for_backward(data, depth, meta)
        ctx.attrs = (grid_size, kernel_size, threshold)
        outs = backprojecting_op.backproject(data, label, depth, meta, label_3d, grid_size, kernel_size, threshold)
        ctx.mark_non_differentiable(outs[1], outs[2])
        return outs

    @staticmethod
    def backward(ctx, grad, _l, _f):
        data, depth, meta = ctx.saved_tensors
        g = backprojecting_op.backproject_grad(data, depth, meta, grad.contiguous(), *ctx.attrs)
        return g, None, None, None, None, None, None, None


def backproject(data, label, depth, meta, label_3d, grid_size, kernel_size, threshold):
    return Backproject.apply(data, label, depth, meta, label_3d, grid_size, kernel_size, threshold)
