"""Three-way parity on the GPU box: the reference's OWN CUDA kernels (oracle/_ref, compiled
unmodified from /root/reference for sm_100a) vs the CPU oracle vs this repo's kernels."""
import numpy as np
import pytest
import torch

from oracle import oracle, ref
from tests.golden import cases
from tests.util import assert_hough_rows_equal, to_np

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def refdev(cuda):
    if not ref.available():
        pytest.skip("oracle/_ref/libposecnn_ref.so not built (needs /root/reference at build time)")
    return cuda


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.mark.parametrize("name", ["h_test_a", "h_train_b", "h_thr_c", "h_skip1_d"])
def test_hough_reference_vs_oracle_vs_ours(refdev, name):
    from posecnn_b200.hough_voting_gpu_layer import hough_voting_gpu_op as op
    sc, is_train, vt, pt, skip = cases.hough_inputs(name)
    d = [T(sc[k], refdev) for k in ("label", "vertex", "extents", "meta")]
    gt = T(sc["gt"], refdev) if len(sc["gt"]) else None
    r_outs, r_n, r_votes = ref.hough_canonical(*d, gt, is_train, vt, pt, skip)
    o_outs, dbg = oracle.hough_voting_gpu(sc["label"], sc["vertex"], sc["extents"], sc["meta"], sc["gt"], is_train, vt, pt,
                                          skip, debug=True)
    # reference kernels vs CPU restatement: votes within the ambiguity bound, rows equal
    dv = np.abs(to_np(r_votes) - dbg["votes"])
    assert (dv <= dbg["ambig"]).all(), f"oracle differs from the reference kernel on {int((dv > dbg['ambig']).sum())} cells"
    assert r_n == dbg["num_rois"]
    assert_hough_rows_equal(r_outs, o_outs, is_train)
    # this repo's kernels vs the reference kernels, directly
    planes = op.hough_vote_planes(*d, skip)
    dv2 = np.abs(to_np(planes) - to_np(r_votes))
    assert (dv2 <= dbg["ambig"]).all(), f"{int((dv2 > dbg['ambig']).sum())} cells differ from the reference kernel"
    got = op.hough_voting_gpu(*d, gt, is_train, vt, pt, skip)
    assert_hough_rows_equal(got, r_outs, is_train)


def test_hough_reference_launcher_as_shipped(refdev):
    """The reference launcher exactly as shipped (atomicAdd list order).  With skip_pixels = 1 every
    pixel votes, so the vote set is order independent and the op output must match ours as a set."""
    from posecnn_b200.hough_voting_gpu_layer import hough_voting_gpu_op as op
    sc, is_train, vt, pt, skip = cases.hough_inputs("h_skip1_d")
    d = [T(sc[k], refdev) for k in ("label", "vertex", "extents", "meta")]
    gt = T(sc["gt"], refdev)
    r_outs, r_n = ref.hough_full(*d, gt, 0, vt, pt, 1)
    got = op.hough_voting_gpu(*d, gt, 0, vt, pt, 1)
    assert r_n == got[0].shape[0]
    assert_hough_rows_equal(got, r_outs, 0, as_set=True)


def test_pixel_ops_reference_vs_ours(refdev):
    from posecnn_b200.average_distance_loss import average_distance_loss_op as aop
    from posecnn_b200.backprojecting_layer import backprojecting_op as bop
    from posecnn_b200.hard_label_layer import hard_label_op as hop
    from posecnn_b200.projecting_layer import projecting_op as pop
    from posecnn_b200.roi_pooling_layer import roi_pooling_op as rop
    dev = refdev
    data, rois, grad = cases.roi_inputs()
    for pc in (0, 1):
        rt, ra = ref.roi_pool(T(data, dev), T(rois, dev), 7, 7, 1.0 / 16.0, pc)
        ot, oa = rop.roi_pool(T(data, dev), T(rois, dev), 7, 7, 1.0 / 16.0, pc)
        np.testing.assert_array_equal(to_np(oa), to_np(ra))
        np.testing.assert_array_equal(to_np(ot), to_np(rt))
        wt_, wa_ = oracle.roi_pool(data, rois, 7, 7, 1.0 / 16.0, pc)
        np.testing.assert_array_equal(wa_, to_np(ra))
    rt, ra = ref.roi_pool(T(data, dev), T(rois, dev), 7, 7, 1.0 / 16.0, 0)
    rg = ref.roi_pool_grad(T(data, dev), T(rois, dev), ra, T(grad, dev), 7, 7, 1.0 / 16.0, 0)
    og = rop.roi_pool_grad(T(data, dev), T(rois, dev), ra, T(grad, dev), 7, 7, 1.0 / 16.0, 0)
    np.testing.assert_allclose(to_np(og), to_np(rg), rtol=1e-5, atol=1e-6)
    prob, gtl = cases.hard_label_inputs()
    for thr in (1.0, 0.5):
        np.testing.assert_array_equal(to_np(hop.hard_label(T(prob, dev), T(gtl, dev), thr)),
                                      to_np(ref.hard_label(T(prob, dev), T(gtl, dev), thr)))
    c = cases.projection_inputs()
    d = {k: T(v, dev) for k, v in c.items()}
    rtd, rtl, rtf = ref.backproject(d["data"], d["label"], d["depth"], d["meta"], d["label_3d"], 16, 3, 0.02)
    otd, otl, otf = bop.backproject(d["data"], d["label"], d["depth"], d["meta"], d["label_3d"], 16, 3, 0.02)
    # same fp32 expressions, same contraction -> identical on the same device
    np.testing.assert_array_equal(to_np(otf), to_np(rtf))
    np.testing.assert_allclose(to_np(otd), to_np(rtd), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(to_np(otl), to_np(rtl), rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(to_np(pop.project(d["vox"], d["depth"], d["meta"], 3, 0.02)),
                                  to_np(ref.project(d["vox"], d["depth"], d["meta"])))
    np.testing.assert_array_equal(to_np(bop.backproject_grad(d["data"], d["depth"], d["meta"], d["g3"], 16)),
                                  to_np(ref.backproject_grad(d["g3"], d["depth"], d["meta"], 48, 64)))
    np.testing.assert_allclose(to_np(pop.project_grad(d["vox"], d["depth"], d["meta"], d["g2"], 3, 0.02)),
                               to_np(ref.project_grad(d["g2"], d["depth"], d["meta"], 16, 3, 0.02)), rtol=1e-6, atol=1e-6)
    pred, targ, wt, pts, sym = cases.avgdist_inputs()
    a = [T(x, dev) for x in (pred, targ, wt, pts, sym)]
    rl, rd = ref.average_distance_loss(*a, 0.01)
    ol, od = aop.average_distance_loss(*a, 0.01)
    np.testing.assert_allclose(to_np(ol), to_np(rl), rtol=1e-4)
    np.testing.assert_allclose(to_np(od), to_np(rd), rtol=1e-3, atol=1e-4 * float(rd.abs().max()))
