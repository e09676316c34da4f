"""Plain PyTorch fp32 restatement of the vgg16_convs graph (lib/networks/vgg16_convs.py:79-212 with the
layer semantics of lib/networks/network.py) — the reference for the network-level parity tests and the network half
of the CPU arm of the full path (`bench.py --impl reference`, oracle/cpu_pipeline.py).
Dense conv2d_transpose with the fixed diagonal bilinear filters, in the reference's op order.

TEST INFRASTRUCTURE / CPU BASELINE ONLY: never imported by the product package."""
import torch
import torch.nn.functional as F

# the reference must be true fp32: no TF32 in cuDNN / cuBLAS
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

# lib/networks/vgg16_convs.py:80-97 (layer name, Cin, Cout | pool)
VGG_CFG = [("conv1_1", 3, 64), ("conv1_2", 64, 64), "pool1", ("conv2_1", 64, 128), ("conv2_2", 128, 128), "pool2",
           ("conv3_1", 128, 256), ("conv3_2", 256, 256), ("conv3_3", 256, 256), "pool3", ("conv4_1", 256, 512),
           ("conv4_2", 512, 512), ("conv4_3", 512, 512), "pool4", ("conv5_1", 512, 512), ("conv5_2", 512, 512),
           ("conv5_3", 512, 512)]


def conv(x, w_hwio, b, relu=True):           # Network.conv, network.py:159-188 (NCHW here)
    y = F.conv2d(x, w_hwio.permute(3, 2, 0, 1), b, padding=w_hwio.shape[0] // 2)
    return F.relu(y) if relu else y


def deconv_filter(k, c):                     # make_deconv_filter, network.py:141-157
    f = (k + 1) // 2
    cc = (2 * f - 1 - f % 2) / (2.0 * f)
    w1 = torch.tensor([1 - abs(x / f - cc) for x in range(k)], dtype=torch.float32)
    bil = w1[:, None] * w1[None, :]
    w = torch.zeros((c, c, k, k))
    for i in range(c):
        w[i, i] = bil
    return w


def deconv(x, k, s):                         # Network.deconv, network.py:207-222 (SAME, out = in * s)
    c = x.shape[1]
    return F.conv_transpose2d(x, deconv_filter(k, c).to(x.device), stride=s, padding=(k - s) // 2)


def trunk(params, x, sfx=""):
    feats = {}
    for item in VGG_CFG:
        if isinstance(item, str):
            x = F.max_pool2d(x, 2)
        else:
            name = item[0]
            x = conv(x, params[f"{name}{sfx}/weights"], params[f"{name}{sfx}/biases"])
            feats[name] = x
    return feats


def heads(params, c4, c5, num_classes):
    """c4, c5: NCHW fp32 (conv4_3, conv5_3).  Returns score [B,C,H,W], label, prob_normalized, vertex_pred."""
    s5 = conv(c5, params["score_conv5/weights"], params["score_conv5/biases"])
    s4 = conv(c4, params["score_conv4/weights"], params["score_conv4/biases"])
    return heads_from_scores(params, s4, s5,
                             conv(c4, params["score_conv4_vertex/weights"], params["score_conv4_vertex/biases"], False),
                             conv(c5, params["score_conv5_vertex/weights"], params["score_conv5_vertex/biases"], False))


def heads_from_scores(params, s4, s5, v4, v5):
    add = s4 + deconv(s5, 4, 2)
    up = deconv(add, 16, 8)
    score = conv(up, params["score/weights"], params["score/biases"])         # 1x1 with ReLU (vgg16_convs.py:141)
    prob = F.softmax(score, dim=1)
    label = torch.argmax(score, dim=1).to(torch.int32)
    addv = v4 + deconv(v5, 4, 2)
    upv = deconv(addv, 16, 8)
    vertex = conv(upv, params["vertex_pred/weights"], params["vertex_pred/biases"], False)
    return score, label, prob, vertex
