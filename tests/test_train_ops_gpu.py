"""GPU parity of the training-side target generation and fused losses against the reference's
own _generate_vertex_targets output (tests/golden/vertex_targets.npz) and the numpy restatements of the TF loss graphs."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle
from tests.golden import cases
from tests.util import to_np

pytestmark = pytest.mark.gpu
T = lambda a, dev: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def test_vertex_targets_match_reference_golden(cuda):
    from posecnn_b200 import train_ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "vertex_targets.npz"))
    label, centers = cases.vertex_target_inputs()
    t, w = train_ops.generate_vertex_targets(T(label, cuda), T(centers, cuda), 10.0)
    np.testing.assert_array_equal(to_np(w), g["weights"])
    got, want = to_np(t), g["targets"]
    # direction components: float64 divide rounded to float32 -- bit exact; log z: double log, <= 1 float32 ulp
    np.testing.assert_array_equal(got[..., 0::3], want[..., 0::3])
    np.testing.assert_array_equal(got[..., 1::3], want[..., 1::3])
    np.testing.assert_allclose(got[..., 2::3], want[..., 2::3], rtol=1.2e-7, atol=0)


def test_vertex_targets_full_size_properties(cuda):
    """640x480x22 at batch 4: unit direction vectors, weights only on labelled pixels of listed classes, equals the oracle."""
    from posecnn_b200 import synth, train_ops
    sc = synth.make_scene(batch=4, height=480, width=640, num_classes=22, seed=99)
    label = sc["label"]
    C = 22
    rng = np.random.default_rng(1)
    centers = np.zeros((4, C, 3), np.float32)
    for b in range(4):
        for c in np.unique(label[b]):
            if c > 0:
                ys, xs = np.where(label[b] == c)
                centers[b, c] = (xs.mean() + rng.normal(), ys.mean() + rng.normal(), rng.uniform(0.5, 1.5))
    t, w = train_ops.generate_vertex_targets(T(label, cuda), T(centers, cuda), 10.0)
    wt, ww = oracle.generate_vertex_targets(label, centers, 10.0)
    np.testing.assert_array_equal(to_np(w), ww)
    np.testing.assert_allclose(to_np(t), wt, rtol=1.2e-7, atol=0)
    tt = to_np(t).reshape(4, 480, 640, C, 3)
    fg = label > 0
    sel = np.take_along_axis(tt, np.maximum(label, 0)[..., None, None].astype(np.int64), axis=3)[..., 0, :]
    n = np.hypot(sel[..., 0], sel[..., 1])[fg]
    assert np.all((np.abs(n - 1) < 1e-5) | (n == 0))


@pytest.mark.parametrize("threshold", [1.0, 0.4])
def test_loss_cross_entropy_hard(cuda, threshold):
    from posecnn_b200 import train_ops
    rng = np.random.default_rng(3)
    B, H, W, C = 2, 48, 64, 22
    logits = rng.standard_normal((B, H, W, C)).astype(np.float32) * 2
    score = logits - np.log(np.exp(logits).sum(3, keepdims=True))          # log-softmax
    prob = np.exp(score).astype(np.float32)
    gt = rng.integers(-1, C, size=(B, H, W)).astype(np.int32)
    want, mask = oracle.loss_cross_entropy_hard(score, prob, gt, threshold)
    loss, count, grad = train_ops.loss_cross_entropy_hard(T(score, cuda), T(prob, cuda), T(gt, cuda), threshold, want_grad=True)
    assert float(count.item()) == mask.sum()                                 # selection: exact
    assert abs(float(loss.item()) - want) <= 1e-5 * abs(want)                # stated tolerance: rel 1e-5 (fp32 result of a double sum)
    np.testing.assert_allclose(to_np(grad), (-mask / (mask.sum() + 1e-10)).astype(np.float32), rtol=1e-6, atol=0)
    # deterministic: bit-identical across launches
    loss2, _ = train_ops.loss_cross_entropy_hard(T(score, cuda), T(prob, cuda), T(gt, cuda), threshold)
    assert float(loss2.item()) == float(loss.item())
    # the fused loss equals the un-fused composition through the Hardlabel op
    from posecnn_b200.hard_label_layer import hard_label_op
    m = hard_label_op.hard_label(T(prob, cuda), T(gt, cuda), threshold)
    comp = -(m.double() * T(score, cuda).double()).sum() / (m.double().sum() + 1e-10)
    assert abs(float(loss.item()) - float(comp.item())) <= 1e-5 * abs(float(comp.item()))


@pytest.mark.parametrize("sigma,n", [(1.0, 4 * 30 * 40 * 66), (3.0, 1001)])
def test_smooth_l1_loss_vertex(cuda, sigma, n):
    from posecnn_b200 import train_ops
    rng = np.random.default_rng(4)
    pred = rng.standard_normal(n).astype(np.float32)
    targ = rng.standard_normal(n).astype(np.float32)
    wgt = np.where(rng.random(n) < 0.2, 10.0, 0.0).astype(np.float32)
    pad = (-n) % 4                                                   # the wrapper needs 16-byte aligned tensors, any length
    want, gwant = oracle.smooth_l1_loss_vertex(pred, targ, wgt, sigma)
    loss, wsum, grad = train_ops.smooth_l1_loss_vertex(T(pred, cuda), T(targ, cuda), T(wgt, cuda), sigma, want_grad=True)
    assert float(wsum.item()) == float(wgt.astype(np.float64).sum())
    assert abs(float(loss.item()) - want) <= 1e-5 * abs(want)
    np.testing.assert_allclose(to_np(grad), gwant, rtol=1e-5, atol=1e-9)
    # gradient check against torch autograd of the reference formula
    p = T(pred, cuda).double().requires_grad_()
    d = T(wgt, cuda).double() * (p - T(targ, cuda).double())
    s2 = sigma ** 2
    sign = (d.abs() < 1.0 / s2).double().detach()
    l = ((d ** 2) * (s2 / 2) * sign + (d.abs() - 0.5 / s2) * (1 - sign)).sum() / (T(wgt, cuda).double().sum() + 1e-10)
    l.backward()
    assert torch.allclose(grad.double(), p.grad, rtol=1e-4, atol=1e-9)
