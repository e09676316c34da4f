"""Deviation of the bf16 tensor-core pose head (RoiPool pair + fc6-fc8 + tanh) from the fp32 torch head on the SAME
trunk features / ROIs, at full size (batch 2 x 480x640, C = 22).  Prints max / mean abs error of poses_tanh."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posecnn_b200 import synth
from posecnn_b200.networks.vgg16_convs import vgg16_convs
from posecnn_b200.roi_pooling_layer import roi_pooling_op as rop
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
B, H, W, C = 2, 480, 640, 22
net = vgg16_convs(num_classes=C, device=dev).init_random(seed=0)
rgb, _ = synth.make_images(B, H, W, seed=3)
data = torch.from_numpy(rgb).to(dev)
meta = torch.from_numpy(np.stack([synth.make_meta(synth.intrinsics(H, W))] * B)).to(dev)
ext = torch.from_numpy(synth.extents_for(C)).to(dev)
net.calibrate_background(data, meta, ext, 0.75)
L = net.forward(data, meta, ext, sync_rois=False)
n = int(L["num_rois"].item())
rois = L["rois_capacity"][:n]
p5, _ = rop.roi_pool(L["conv5_3"], rois, 7, 7, 1 / 16.0, 0)
p4, _ = rop.roi_pool(L["conv4_3"], rois, 7, 7, 1 / 8.0, 0)
x = (p5 + p4).reshape(n, -1)
P = net.params
h6 = torch.relu(x @ P["fc6/weights"] + P["fc6/biases"])
h7 = torch.relu(h6 @ P["fc7/weights"] + P["fc7/biases"])
pre = h7 @ P["fc8/weights"] + P["fc8/biases"]
want = torch.tanh(pre)
got = L["poses_tanh"][:n]
e = (got - want).abs()
print("rois", n, "| pre-tanh |x| max %.3f" % pre.abs().max().item(), "| poses_tanh abs err max %.3e mean %.3e" % (e.max().item(), e.mean().item()),
      "| fc6 rel-L2 %.3e" % ((L["fc6"][:n, :4096].float() - h6).norm() / h6.norm()).item())
