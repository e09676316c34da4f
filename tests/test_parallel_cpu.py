"""Multi-process path on CPU (gloo, world_size 2): record packing and the all-gather of per-rank
pose-hypothesis records (posecnn_b200/parallel.py).  The GPU path uses the same code over NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from posecnn_b200 import parallel


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _fake_layers(rank, C, cap_rows, n_valid):
    g = torch.Generator().manual_seed(100 + rank)
    rois = torch.zeros((cap_rows, 7))
    rois[:n_valid, 0] = torch.arange(n_valid) % 4                      # local batch index
    rois[:n_valid, 1] = torch.randint(1, C, (n_valid,), generator=g).float()
    rois[:n_valid, 2:] = torch.rand((n_valid, 5), generator=g) * 100
    return dict(rois_capacity=rois, num_rois=torch.tensor([n_valid], dtype=torch.int32),
                poses_init=torch.rand((cap_rows, 7), generator=g), poses_tanh=torch.rand((cap_rows, 4 * C), generator=g))


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    C, cap, local_batch = 6, 16, 4
    L = _fake_layers(rank, C, cap, n_valid=5 + 3 * rank)
    rec = parallel.pack_records(L, C, rank * local_batch)
    allrec = parallel.all_gather_records(rec, world)
    # post-NMS detections (final payload) and the training-side loss normaliser
    nd = 2 + rank
    L.update(detections_rois=L["rois_capacity"].clone(), detections_poses=L["poses_init"].clone(),
             num_detections=torch.tensor([nd], dtype=torch.int32))
    det = parallel.all_gather_records(parallel.pack_detections(L, rank * local_batch), world)
    # the per-step pipeline (double-buffered copy + collective; CPU tensors take the stream-less branch of the same protocol)
    pipe = parallel.GatherPipeline(world, rec)
    piped = []
    for step in range(3):
        pipe.before_step()
        i = pipe.submit(rec + float(step))
        piped.append(pipe.results(i).clone())
    pipe.drain()
    loss, gscale = parallel.global_mean_loss(torch.tensor(1.0 + rank), 3 + 5 * rank, world)
    torch.save(dict(rec=rec, allrec=allrec, det=det, loss=loss, gscale=gscale, piped=piped), os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_all_gather_records_gloo_world2(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    C, cap, local_batch = 6, 16, 4
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    assert torch.equal(outs[0]["allrec"], outs[1]["allrec"])           # every rank holds the same gathered table
    allrec = outs[0]["allrec"]
    assert allrec.shape == (world * cap, parallel.record_width(C))
    for r in range(world):
        blk = allrec[r * cap:(r + 1) * cap]
        assert torch.equal(blk, outs[r]["rec"])
        n_valid = 5 + 3 * r
        assert blk[:, -1].sum().item() == n_valid and not blk[n_valid:, :14].any()
        L = _fake_layers(r, C, cap, n_valid)
        # batch indices are made global: rank * local_batch + local index
        np.testing.assert_array_equal(blk[:n_valid, 0].numpy(), (L["rois_capacity"][:n_valid, 0] + r * local_batch).numpy())
        assert torch.equal(blk[:n_valid, 1:7], L["rois_capacity"][:n_valid, 1:7])
        assert torch.equal(blk[:n_valid, 14:14 + 4 * C], L["poses_tanh"][:n_valid])
        det = outs[0]["det"][r * cap:(r + 1) * cap]
        assert det.shape[1] == parallel.detection_width() and det[:, -1].sum().item() == 2 + r and not det[2 + r:].any()
        assert torch.equal(det[:2 + r, 7:14], L["poses_init"][:2 + r])
        np.testing.assert_array_equal(det[:2 + r, 0].numpy(), (L["rois_capacity"][:2 + r, 0] + r * local_batch).numpy())
    assert torch.equal(outs[0]["det"], outs[1]["det"])
    for step in range(3):   # GatherPipeline: every rank sees [rank 0 records | rank 1 records] of that step
        want = torch.cat([outs[0]["rec"] + float(step), outs[1]["rec"] + float(step)])
        assert torch.equal(outs[0]["piped"][step], want) and torch.equal(outs[1]["piped"][step], want)
    # ranks hold means 1.0 (3 rows) and 2.0 (8 rows): global mean = (3 + 16) / 11; grad scales N_r * world / N
    assert abs(outs[0]["loss"].item() - 19.0 / 11.0) < 1e-6 and abs(outs[1]["loss"].item() - 19.0 / 11.0) < 1e-6
    assert abs(outs[0]["gscale"].item() - 6.0 / 11.0) < 1e-6 and abs(outs[1]["gscale"].item() - 16.0 / 11.0) < 1e-6


def test_pack_records_single_process():
    L = _fake_layers(0, 4, 8, 3)
    rec = parallel.pack_records(L, 4, batch_offset=64)
    assert rec.shape == (8, parallel.record_width(4))
    assert rec[:3, 0].tolist() == [64.0, 65.0, 66.0] and rec[3:, 0].abs().sum() == 0
    assert parallel.all_gather_records(rec, 1) is rec


def test_shard_range_and_capacity():
    """contiguous shards of ONE global batch; ROI budget from the global batch size."""
    for B, G in ((32, 1), (32, 2), (32, 4), (32, 8), (10, 4), (3, 8)):
        parts = [parallel.shard_range(B, r, G) for r in range(G)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == B
        for (o0, c0), (o1, _) in zip(parts, parts[1:]):
            assert o1 == o0 + c0
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    assert parallel.roi_capacity(32, 32) == 128 and parallel.roi_capacity(32, 4) == 16    # 128 // 32 = 4 maxima per image
    assert parallel.roi_capacity(64, 8, is_train=True) == 2 * 8 * 9 and parallel.roi_capacity(256, 32) == 1
    table = torch.zeros((6, parallel.detection_width()))
    table[[0, 1, 3], -1] = 1.0
    assert parallel.compact_records(table).shape[0] == 3
