"""Gradient wiring of Averagedistance — mirrors lib/average_distance_loss/average_distance_loss_op_grad.py:5-14
(gradient for the prediction only = upstream * bottom_diff)."""
import torch

from . import average_distance_loss_op


class AverageDistance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prediction, target, weight, point, symmetry, margin):
        loss, diff = average_distance_loss_op.average_distance_loss(prediction, target, weight, point, symmetry, margin)
        ctx.save_for_backward(diff)
        ctx.mark_non_differentiable(diff)
        return loss, diff

    @staticmethod
    def backward(ctx, grad, _):
        (diff,) = ctx.saved_tensors
        g = average_distance_loss_op.average_distance_loss_grad(diff, grad.contiguous().reshape(-1)[:1])
        return g, None, None, None, None, None


def average_distance_loss(prediction, target, weight, point, symmetry, margin):
    return AverageDistance.apply(prediction, target, weight, point, symmetry, margin)
