"""Gradient wiring of Project — mirrors lib/projecting_layer/projecting_op_grad.py."""
import torch

from . import projecting_op


class Project(torch.autograd.Function):
    @staticmethod
    def forward(ctx, data, depth, meta, kernel_size, threshold):
        ctx.save_for_backward(data, depth, meta)
        ctx.attrs = (kernel_size, threshold)
        return projecting_op.project(data, depth, meta, kernel_size, threshold)

    @staticmethod
    def backward(ctx, grad):
        data, depth, meta = ctx.saved_tensors
        g = projecting_op.project_grad(data, depth, meta, grad.contiguous(), *ctx.attrs)
        return g, None, None, None, None


def project(data, depth, meta, kernel_size, threshold):
    return Project.apply(data, depth, meta, kernel_size, threshold)
