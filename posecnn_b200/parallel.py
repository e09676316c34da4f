"""Image-sharded multi-GPU inference: one process per GPU, every rank runs the whole
network on its own contiguous shard of images; the only exchange is one NCCL all-gather of the fixed-size
per-rank hypothesis records (ROI rows, initial poses, regressed quaternions, row count) over NVLink.

The reference has no distributed code at all (SURVEY finding 9); this is new design.  Each rank's shard is
treated as one reference batch: the MAX_ROI / batch_size cap (hough_voting_gpu_op.cu.cc:733) is applied to
the rank-local batch, and batch indices in the gathered ROI rows are made global (rank * local_batch + b).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def record_width(num_classes: int) -> int:
    return 7 + 7 + 4 * num_classes + 1  # roi row, initial pose, tanh quaternions, valid flag


def pack_records(layers: dict, num_classes: int, rank: int, local_batch: int) -> torch.Tensor:
    """[cap_rows, 7 + 7 + 4C + 1] f32; rows >= num_rois carry valid = 0 (no host sync needed)."""
    rois = layers["rois_capacity"]
    n = rois.shape[0]
    rec = torch.zeros((n, record_width(num_classes)), dtype=torch.float32, device=rois.device)
    rec[:, 0:7] = rois
    rec[:, 0] += float(rank * local_batch)
    rec[:, 7:14] = layers["poses_init"][:n]
    if "poses_tanh" in layers:
        rec[:, 14:14 + 4 * num_classes] = layers["poses_tanh"][:n]
    valid = torch.arange(n, device=rois.device) < layers["num_rois"].to(torch.int64)
    rec[:, -1] = 1.0
    rec *= valid.float()[:, None]  # rows beyond num_rois are all-zero (valid flag 0)
    return rec


def detection_width() -> int:
    return 7 + 7 + 1  # roi row, final pose [quaternion | translation], valid flag


def pack_detections(layers: dict, rank: int, local_batch: int) -> torch.Tensor:
    """Post-NMS records (the all-gather payload is final): [cap_rows, 15] f32 =
    [batch(global), cls, x1, y1, x2, y2, score | qw, qx, qy, qz, tx, ty, tz | valid]; rows >= num_detections are zero."""
    rois, poses, n = layers["detections_rois"], layers["detections_poses"], layers["num_detections"]
    cap = rois.shape[0]
    valid = (torch.arange(cap, device=rois.device) < n.to(torch.int64)).float()[:, None]
    shift = torch.zeros((1, 7), dtype=torch.float32, device=rois.device)
    shift[0, 0] = float(rank * local_batch)
    return torch.cat([(rois + shift) * valid, poses * valid, valid], 1)


def all_gather_records(rec: torch.Tensor, world: int) -> torch.Tensor:
    """[world * cap_rows, width]; latency-bound (tens of KB), one ncclAllGather."""
    if world == 1:
        return rec
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous())
    return out


def global_mean_loss(loss_local: torch.Tensor, n_local: int | torch.Tensor, world: int):
    """Training-side normaliser: a loss that the op divides by the LOCAL row count
    (Averagedistance: sum / (2 N P), average_distance_loss_op_gpu.cu.cc:181,196) becomes the single-GPU value when every
    rank's mean is weighted by its row count: L = sum_r L_r N_r / sum_r N_r.  One all-reduce of two scalars.
    Returns (global loss, grad_scale) where grad_scale = world * N_r / sum N multiplies the local bottom_diff so that the
    usual gradient all-reduce AVERAGE over ranks reproduces d L / d theta."""
    n = torch.as_tensor(n_local, dtype=torch.float32, device=loss_local.device).reshape(())
    pair = torch.stack([loss_local.reshape(()).float() * n, n])
    if world > 1:
        dist.all_reduce(pair, op=dist.ReduceOp.SUM)
    total = pair[1].clamp(min=1.0)
    return pair[0] / total, n * world / total
