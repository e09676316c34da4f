"""Gradient wiring of Hardlabel — mirrors lib/hard_label_layer/hard_label_op_grad.py (zeros)."""
import torch

from . import hard_label_op


class HardLabel(torch.autograd.Function):
    @staticmethod
    def forward(ctx, prob, gt, threshold):
        ctx.save_for_backward(prob, gt)
        return hard_label_op.hard_label(prob, gt, threshold)

    @staticmethod
    def backward(ctx, grad):
        prob, gt = ctx.saved_tensors
        g_prob, _ = hard_label_op.hard_label_grad(prob, gt, grad)
        return g_prob, None, None


def hard_label(prob, gt, threshold):
    return HardLabel.apply(prob, gt, threshold)
