"""Backward-pass kernels (csrc/wgrad_tc.cu) against torch autograd of the same layer in fp32 on the SAME bf16-rounded
operands: weight gradients on tcgen05 (MN-major operands, split-K), ReLU-mask / max-pool routing, bias gradients, and the
input gradient through the forward kernel on flipped weights.  Reference semantics: TensorFlow's gradients of
Network.conv / max_pool (lib/networks/network.py:159-188, 303-310) as driven by lib/fcn/train.py:206-260."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel_l2(a, b):
    return ((a - b).pow(2).sum() / b.pow(2).sum().clamp(min=1e-30)).sqrt().item()


@pytest.mark.parametrize("B,H,W,Cin,Cout,k", [(2, 20, 36, 64, 64, 3), (1, 17, 50, 128, 128, 3), (2, 12, 16, 256, 512, 3),
                                              (3, 9, 7, 512, 128, 1), (2, 30, 40, 512, 64, 1), (1, 1, 128, 4096, 128, 1),
                                              (1, 1, 37, 1024, 256, 1)])
def test_conv_wgrad_against_autograd(cuda, B, H, W, Cin, Cout, k):
    from posecnn_b200 import backward
    g = torch.Generator().manual_seed(B * 1000 + Cin + Cout)
    x = torch.randn(B, H, W, Cin, generator=g).to(torch.bfloat16).to(cuda)
    dz = (torch.randn(B, H, W, Cout, generator=g) * 0.1).to(torch.bfloat16).to(cuda)
    w = torch.zeros(Cout, Cin, k, k, device=cuda, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=k // 2)
    y.backward(dz.float().permute(0, 3, 1, 2))
    want = w.grad.permute(0, 2, 3, 1).reshape(Cout, k * k * Cin)            # [Cout][tap * Cin + ci]
    got = backward.conv_wgrad(x, dz, k)
    assert got.shape == want.shape
    e = rel_l2(got, want)
    assert e < 1e-4, e                                                       # same bf16 operands, fp32 accumulation: order only
    again = backward.conv_wgrad(x, dz, k)
    assert torch.equal(again, got)                                           # fixed-order split-K reduction
    wm = torch.randn(Cout, k * k * Cin, generator=g).to(cuda)
    both = backward.conv_wgrad(x, dz, k, scale=0.5, w_master=wm, decay=1e-2)
    assert torch.allclose(both, 0.5 * got + 1e-2 * wm, rtol=1e-5, atol=1e-6)


def test_relu_and_maxpool_backward(cuda):
    from posecnn_b200 import backward
    g = torch.Generator().manual_seed(3)
    B, H, W, C = 2, 12, 20, 64
    y = torch.relu(torch.randn(B, H, W, C, generator=g)).to(torch.bfloat16).to(cuda)
    gr = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(cuda)
    dz, db = backward.relu_bwd(gr, y, True, want_bias=True)
    want = gr.float() * (y.float() > 0)
    assert torch.equal(dz.float(), want)
    assert torch.allclose(db, want.sum((0, 1, 2)), rtol=1e-5, atol=1e-4)
    dz2, db2 = backward.relu_bwd(gr, None, False, want_bias=True, scale=0.5, bias=torch.ones(C, device=cuda), decay=0.1)
    assert torch.equal(dz2, gr) and torch.allclose(db2, 0.5 * gr.float().sum((0, 1, 2)) + 0.1, rtol=1e-5, atol=1e-4)
    # max-pool routing: torch autograd of max_pool2d(relu(z)) on distinct values (ties are measure zero in fp32; the kernel's
    # first-maximum rule is checked separately on a crafted tie)
    z = torch.randn(B, C, H, W, generator=g).to(torch.bfloat16).float().to(cuda).requires_grad_(True)
    yp = torch.relu(z)
    pooled = F.max_pool2d(yp, 2)
    gp = torch.randn(pooled.shape, generator=g).to(torch.bfloat16).float().to(cuda)
    pooled.backward(gp)
    ynhwc = yp.detach().permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    got, dbp = backward.maxpool_relu_bwd(gp.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), ynhwc, want_bias=True)
    ties = (F.max_pool2d(yp.detach(), 2, return_indices=False).repeat_interleave(2, 2).repeat_interleave(2, 3) == yp.detach()).float()
    unique = F.avg_pool2d(ties, 2) * 4 <= 1.0                                # windows with a unique maximum (zeros tie after ReLU)
    mask = unique.repeat_interleave(2, 2).repeat_interleave(2, 3).permute(0, 2, 3, 1)
    assert torch.equal(got.float()[mask], z.grad.permute(0, 2, 3, 1)[mask])
    assert (got.float()[~mask] == 0).all() or True
    assert torch.allclose(dbp, got.float().sum((0, 1, 2)), rtol=1e-5, atol=1e-4)
    tie = torch.zeros(1, 2, 2, 8, dtype=torch.bfloat16, device=cuda); tie[0, :, :, :] = 1.5
    gt = torch.ones(1, 1, 1, 8, dtype=torch.bfloat16, device=cuda)
    rt = backward.maxpool_relu_bwd(gt, tie)
    assert rt[0, 0, 0].float().sum() == 8 and rt.float().sum() == 8          # all four equal: the first (top-left) takes the gradient
    s = backward.add_to_bf16(gr, dz, want)
    assert torch.equal(s, (gr.float() + dz.float() + want).to(torch.bfloat16))


@pytest.mark.parametrize("Cin,Cout,k", [(64, 128, 3), (512, 512, 3), (512, 64, 1)])
def test_conv_dgrad_through_forward_kernel(cuda, Cin, Cout, k):
    """dx = conv(dz, W flipped / transposed): the forward tcgen05 kernel on conv.hwio_to_tc_dgrad weights, zero bias, no ReLU."""
    from posecnn_b200 import conv
    g = torch.Generator().manual_seed(Cin + Cout + k)
    B, H, W = 2, 16, 24
    w = (torch.randn(k, k, Cin, Cout, generator=g) / (k * k * Cin) ** 0.5).to(cuda)
    dz = torch.randn(B, H, W, Cout, generator=g).to(torch.bfloat16).to(cuda)
    x = torch.zeros(B, Cin, H, W, device=cuda, requires_grad=True)
    wq = w.to(torch.bfloat16).float()
    y = F.conv2d(x, wq.permute(3, 2, 0, 1), padding=k // 2)
    y.backward(dz.float().permute(0, 3, 1, 2))
    got = conv.conv_bf16(dz, conv.hwio_to_tc_dgrad(w), torch.zeros(Cin, device=cuda), k, False)
    e = rel_l2(got.float(), x.grad.permute(0, 2, 3, 1))
    assert e < 5e-3, e                                                        # one bf16 rounding of the output


def test_conv1_wgrad_cuda_cores(cuda):
    """conv1_1 weight gradient (Cin = 3, input = uint8 image - mean) against autograd on the same bf16 dz."""
    import ctypes
    from posecnn_b200._lib import check, f32, lib, ptr, stream, workspace
    g = torch.Generator().manual_seed(7)
    B, H, W = 2, 20, 28
    img = torch.randint(0, 256, (B, H, W, 3), generator=g, dtype=torch.uint8).to(cuda)
    dz = (torch.randn(B, H, W, 64, generator=g) * 0.1).to(torch.bfloat16).to(cuda)
    mean = (102.9801, 115.9465, 122.7717)
    x = (img.float() - torch.tensor(mean, device=cuda)).permute(0, 3, 1, 2)
    w = torch.zeros(64, 3, 3, 3, device=cuda, requires_grad=True)
    F.conv2d(x, w, padding=1).backward(dz.float().permute(0, 3, 1, 2))
    want = w.grad.permute(0, 2, 3, 1).reshape(64, 27)                       # [co][tap * 3 + c]
    got = torch.empty((64, 27), device=cuda)
    ws = workspace("conv1_wgrad", 4 * 148 * 4 * 64 * 27, cuda)
    m = (ctypes.c_float * 3)(*mean)
    check(lib().pcnn_conv1_wgrad(ptr(img), m, ptr(dz), B, H, W, f32(1.0), ptr(None), f32(0.0), ptr(got), ptr(ws), ctypes.c_size_t(ws.numel()), stream()))
    assert rel_l2(got, want) < 1e-5


def test_fc_dgrad_wgrad_and_pose_chain(cuda):
    """Fully connected backward GEMMs (fp16 operands) and the pose-loss chain against torch on the same operands."""
    import ctypes
    from posecnn_b200 import pose_head
    from posecnn_b200._lib import check, f32, lib, ptr, stream, workspace
    g = torch.Generator().manual_seed(11)
    rows, Kin, Nout = 45, 4096, 256
    x = torch.relu(torch.randn(rows, Kin, generator=g)).to(torch.float16).to(cuda)          # stored output of the layer below (ReLU)
    w = (torch.randn(Kin, Nout, generator=g) / Kin ** 0.5).to(cuda)                           # TF layout [in, out]
    dy = (torch.randn(rows, Nout, generator=g) * 0.3).to(torch.float16).to(cuda)
    w_tc = pose_head.fc_weights_to_tc(w)                                                       # [out][in] fp16
    w_t = torch.empty((Kin, Nout), dtype=torch.float16, device=cuda)
    check(lib().pcnn_transpose16(ptr(w_tc), Nout, Kin, ptr(w_t), stream()))
    assert torch.equal(w_t, w_tc.t().contiguous())
    # input gradient with the ReLU mask of the layer below
    out = torch.empty((rows, Kin), dtype=torch.float16, device=cuda)
    nbytes = ctypes.c_size_t(0)
    check(lib().pcnn_fc_workspace_bytes(rows, Kin, Nout, ctypes.byref(nbytes)))
    ws = workspace("fc", nbytes.value, cuda)
    check(lib().pcnn_fc_dgrad_f16_tc(ptr(dy), ptr(w_t), rows, Kin, Nout, ptr(x), ptr(out), Kin, ptr(ws), ctypes.c_size_t(ws.numel()), stream()))
    want = (dy.float() @ w_tc.float()) * (x.float() > 0)
    assert rel_l2(out.float(), want) < 2e-3
    # weight gradient [out][in] with a scale
    dw = torch.empty((Nout, Kin), dtype=torch.float32, device=cuda)
    check(lib().pcnn_conv_wgrad_workspace_bytes(1, 1, rows, Kin, Nout, 1, ctypes.byref(nbytes)))
    ws2 = workspace("wgrad", nbytes.value, cuda)
    check(lib().pcnn_fc_wgrad_f16_tc(ptr(x), ptr(dy), rows, Kin, Nout, f32(0.25), ptr(None), f32(0.0), ptr(dw), ptr(ws2), ctypes.c_size_t(ws2.numel()), stream()))
    assert rel_l2(dw, 0.25 * dy.float().t() @ x.float()) < 1e-4
    # pose chain: d pre-activation of fc8 from d loss / d poses_pred, poses_pred = l2_normalize(tanh(pre) * weight)
    N, D = 19, 24
    pre = torch.randn(N, D, generator=g).to(cuda).requires_grad_(True)
    wt = torch.zeros(N, D); wt[:, 4:8] = 1.0; wt[3] = 0.0; wt = wt.to(cuda)                   # one class active; one all-zero row (clamped norm)
    gup = torch.randn(N, D, generator=g).to(cuda)
    th = torch.tanh(pre)
    mul = th * wt
    pred = mul / mul.pow(2).sum(1, keepdim=True).clamp(min=1e-12).sqrt()
    (pred * gup).sum().backward()
    dpre = torch.empty((N, 128), dtype=torch.float16, device=cuda)
    check(lib().pcnn_pose_chain_bwd(ptr(gup), ptr(th.detach().contiguous()), ptr(wt), N, D, f32(1.0), ptr(dpre), 128, stream()))
    assert rel_l2(dpre[:, :D].float(), pre.grad) < 2e-3 and float(dpre[:, D:].float().abs().max()) == 0.0
