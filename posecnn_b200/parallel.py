"""Image-sharded multi-GPU inference: one process per GPU; ONE global batch is cut into contiguous
image shards (rank r of G holds images [r B/G, (r+1) B/G)); weights, extents and points are replicated.  Every rank runs
the whole network on its shard with the reference's ROI budget of the GLOBAL batch (index_size = 128 / B_global,
hough_voting_gpu_op.cu.cc:733) and GLOBAL batch indices in the ROI rows, so that the rank-order concatenation of the
per-rank rows is exactly the single-GPU result.  The only exchange is one NCCL all-gather per step of the fixed-size
post-NMS pose-hypothesis records over NVLink; it is issued on a communication stream from a double-buffered copy of
the step's records, so the next step's CUDA graph replays while the previous step's records travel.

The reference has no distributed code at all (SURVEY finding 9); this is new design.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(global_batch: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous shard of a global batch: (first image, number of images); the first `global_batch % world` ranks
    hold one image more."""
    base, rem = divmod(int(global_batch), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def roi_capacity(global_batch: int, local_batch: int, is_train: bool = False) -> int:
    """Rows a shard can emit: (128 // B_global) maxima per image (x9 jitter rows in train mode), at least one row."""
    return max(1, (128 // int(global_batch)) * int(local_batch) * (9 if is_train else 1))


def record_width(num_classes: int) -> int:
    return 7 + 7 + 4 * num_classes + 1  # roi row, initial pose, tanh quaternions, valid flag


def pack_records(layers: dict, num_classes: int, batch_offset: int = 0) -> torch.Tensor:
    """Pre-NMS hypothesis records [cap_rows, 7 + 7 + 4C + 1] f32 (ROI row, initial pose, regressed quaternions, valid);
    rows >= num_rois are zero (no host sync needed).  `batch_offset` is added to the batch column for callers whose ROI
    rows still carry shard-local indices (forward(batch_offset=...) already writes global ones)."""
    rois = layers["rois_capacity"]
    n = rois.shape[0]
    rec = torch.zeros((n, record_width(num_classes)), dtype=torch.float32, device=rois.device)
    rec[:, 0:7] = rois
    rec[:, 0] += float(batch_offset)
    rec[:, 7:14] = layers["poses_init"][:n]
    if "poses_tanh" in layers:
        rec[:, 14:14 + 4 * num_classes] = layers["poses_tanh"][:n]
    valid = torch.arange(n, device=rois.device) < layers["num_rois"].to(torch.int64)
    rec[:, -1] = 1.0
    rec *= valid.float()[:, None]  # rows beyond num_rois are all-zero (valid flag 0)
    return rec


def detection_width() -> int:
    return 7 + 7 + 1  # roi row, final pose [quaternion | translation], valid flag


def pack_detections(layers: dict, batch_offset: int = 0) -> torch.Tensor:
    """Post-NMS records (the all-gather payload is final): [cap_rows, 15] f32 =
    [batch(global), cls, x1, y1, x2, y2, score | qw, qx, qy, qz, tx, ty, tz | valid]; rows >= num_detections are zero.
    Static shapes, no host sync: capturable inside the step's CUDA graph."""
    rois, poses, n = layers["detections_rois"], layers["detections_poses"], layers["num_detections"]
    cap = rois.shape[0]
    valid = (torch.arange(cap, device=rois.device) < n.to(torch.int64)).float()[:, None]
    if batch_offset:
        shift = torch.zeros((1, 7), dtype=torch.float32, device=rois.device)
        shift[0, 0] = float(batch_offset)
        rois = rois + shift
    return torch.cat([rois * valid, poses * valid, valid], 1)


def all_gather_records(rec: torch.Tensor, world: int) -> torch.Tensor:
    """[world * cap_rows, width]; latency-bound (a few KB per rank), one ncclAllGather.  Every rank must pass the same
    shape (use roi_capacity of the LARGEST shard when the batch does not divide evenly)."""
    if world == 1:
        return rec
    out = torch.empty((world * rec.shape[0], rec.shape[1]), dtype=rec.dtype, device=rec.device)
    dist.all_gather_into_tensor(out, rec.contiguous())
    return out


def compact_records(gathered: torch.Tensor) -> torch.Tensor:
    """Valid rows of a gathered record table in rank order = the rows a single GPU would have produced for the whole batch."""
    return gathered[gathered[:, -1] > 0]


class GatherPipeline:
    """Per-step all-gather of the records on a communication stream, double-buffered against the compute stream.

        step i (compute stream):  graph.replay()  ->  records_i (a static buffer the next replay overwrites)
        comm stream:              wait(step i done) -> copy records_i to slot[i & 1] -> [copied_i] -> all-gather -> out[i & 1]
        step i + 1:               waits only for copied_i (a few-KB device copy), not for the collective

    so a rank never stalls on a slower peer inside a step; the collective of step i overlaps the compute of step i + 1.
    `results(i)` hands back the gathered table of step i after making the caller's stream wait for it."""

    def __init__(self, world: int, like: torch.Tensor):
        self.world = world
        dev = like.device
        self.comm = torch.cuda.Stream(device=dev) if like.is_cuda else None
        self.slot = [torch.empty_like(like) for _ in range(2)]
        self.out = [torch.empty((world * like.shape[0], like.shape[1]), dtype=like.dtype, device=dev) for _ in range(2)]
        self.copied = [torch.cuda.Event() for _ in range(2)] if like.is_cuda else None
        self.done = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if like.is_cuda else None
        self.start = [torch.cuda.Event(enable_timing=True) for _ in range(2)] if like.is_cuda else None
        self.step = 0

    def before_step(self):
        """Call on the compute stream before the replay that overwrites the static record buffer."""
        if self.comm is not None and self.step > 0:
            torch.cuda.current_stream().wait_event(self.copied[(self.step - 1) & 1])

    def submit(self, records: torch.Tensor) -> int:
        """Call on the compute stream right after the step's graph replay."""
        i = self.step
        b = i & 1
        if self.comm is None:       # CPU / gloo path (tests): same protocol without streams
            self.slot[b].copy_(records)
            if self.world > 1:
                dist.all_gather_into_tensor(self.out[b], self.slot[b])
            else:
                self.out[b].copy_(self.slot[b])
        else:
            self.comm.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self.comm):
                self.slot[b].copy_(records, non_blocking=True)
                self.copied[b].record(self.comm)
                self.start[b].record(self.comm)
                if self.world > 1:
                    dist.all_gather_into_tensor(self.out[b], self.slot[b])
                else:
                    self.out[b].copy_(self.slot[b], non_blocking=True)
                self.done[b].record(self.comm)
        self.step += 1
        return i

    def results(self, i: int) -> torch.Tensor:
        b = i & 1
        if self.comm is not None:
            torch.cuda.current_stream().wait_event(self.done[b])
        return self.out[b]

    def last_gather_ms(self, i: int) -> float:
        """Device time of step i's collective on the comm stream (includes waiting for the slowest rank); call after a
        synchronize."""
        b = i & 1
        return self.start[b].elapsed_time(self.done[b])

    def drain(self):
        if self.comm is not None:
            torch.cuda.current_stream().wait_stream(self.comm)


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
