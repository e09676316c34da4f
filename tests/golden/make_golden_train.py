"""Generate tests/golden/vertex_targets.npz by EXECUTING THE REFERENCE'S OWN `_generate_vertex_targets`
(lib/gt_synthesize_layer/minibatch.py:543-602).  The module is Python 2 and cannot be imported, so the function's
source text is cut out of the file unmodified and exec'd with the three names it needs from its module scope
(`np`, `math`, `cfg`) plus `xrange = range`.

    python tests/golden/make_golden_train.py        # needs /root/reference; the .npz is committed
"""
import math
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.golden import cases  # noqa: E402

W_INSIDE = 10.0   # cfg.TRAIN.VERTEX_W_INSIDE (experiments/cfgs/lov_color_2d.yml)


def reference_function():
    src = open("/root/reference/lib/gt_synthesize_layer/minibatch.py").read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("def _generate_vertex_targets"))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith("def "))
    cfg = types.SimpleNamespace(TRAIN=types.SimpleNamespace(VERTEX_REG_2D=True, VERTEX_REG_3D=False, VERTEX_W_INSIDE=W_INSIDE))
    ns = dict(np=np, math=math, cfg=cfg, xrange=range)
    exec("\n".join(src[start:end]), ns)
    return ns["_generate_vertex_targets"]


def main():
    fn = reference_function()
    label, centers = cases.vertex_target_inputs()
    B, H, W = label.shape
    C = centers.shape[1]
    targets = np.zeros((B, H, W, 3 * C), np.float32)
    weights = np.zeros((B, H, W, 3 * C), np.float32)
    for b in range(B):
        listed = np.array([c for c in range(1, C) if centers[b, c, 2] > 0], dtype=np.float32)   # cls_indexes of the meta data
        center = np.stack([centers[b, int(c), :2] for c in listed]).astype(np.float32)          # [n, 2]
        poses = np.zeros((3, 4, len(listed)), np.float32)
        poses[2, 3, :] = [centers[b, int(c), 2] for c in listed]
        fn(label[b], listed, center, poses, C, None, None, None, False, None, targets[b], weights[b])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vertex_targets.npz"), targets=targets, weights=weights)
    print("targets nonzero", int((targets != 0).sum()), "weights nonzero", int((weights != 0).sum()))
    # multi-instance branch (minibatch.py:549-573): same function, is_multi_instances = 1
    label, mask, inst = cases.vertex_target_multi_inputs()
    B, H, W = label.shape
    targets = np.zeros((B, H, W, 3 * C), np.float32)
    weights = np.zeros((B, H, W, 3 * C), np.float32)
    for b in range(B):
        rows = [r for r in inst[b] if r[4] > 0]
        cls_indexes = np.array([r[0] for r in rows], dtype=np.float32)
        cls_indexes_old = np.array([r[1] - 1 for r in rows], dtype=np.float32)
        center = np.array([[r[2], r[3]] for r in rows], dtype=np.float32)
        poses = np.zeros((3, 4, len(rows)), np.float32)
        poses[2, 3, :] = [r[4] for r in rows]
        fn(label[b], cls_indexes, center, poses, C, None, None, mask[b], True, cls_indexes_old, targets[b], weights[b])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "vertex_targets_multi.npz"), targets=targets, weights=weights)
    print("multi-instance: targets nonzero", int((targets != 0).sum()), "weights nonzero", int((weights != 0).sum()))


if __name__ == "__main__":
    main()
