"""Debug: whole batch vs per-image shards, layer by layer."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posecnn_b200 import parallel, synth
from posecnn_b200.networks.vgg16_convs import vgg16_convs
dev = torch.device("cuda:0")
B, H, W, C = 4, 96, 128, 6
net = vgg16_convs(num_classes=C, device=dev).init_random(seed=0, bias_std=0.05)
rgb, _ = synth.make_images(B, H, W, seed=9)
data = torch.from_numpy(rgb).to(dev)
meta = torch.from_numpy(np.stack([synth.make_meta(synth.intrinsics(H, W))] * B)).to(dev)
ext = torch.from_numpy(synth.extents_for(C)).to(dev)
for dense in (True, False):
    Lw = dict(net.forward(data, meta, ext, sync_rois=False, dense_vertex=dense)); low_w = net._last_lowres.clone()
    Lw = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in Lw.items()}
    for i in range(B):
        Ls = net.forward(data[i:i + 1], meta[i:i + 1], ext, sync_rois=False, dense_vertex=dense, batch_global=B, batch_offset=i)
        low_s = net._last_lowres
        msg = []
        for k in ("conv4_3", "conv5_3", "label_2d") + (("vertex_pred",) if dense else ()):
            msg.append("%s %s" % (k, bool(torch.equal(Lw[k][i:i + 1], Ls[k]))))
        msg.append("lowres %s" % bool(torch.equal(low_w[i:i + 1], low_s)))
        n = int(Ls["num_rois"].item())
        rw = Lw["rois_capacity"][: int(Lw["num_rois"].item())]
        sel = rw[:, 0] == i
        msg.append("rois %s" % bool(torch.equal(rw[sel], Ls["rois_capacity"][:n])))
        pw = Lw["poses_init"][: int(Lw["num_rois"].item())][sel]
        msg.append("poses_init %s" % bool(torch.equal(pw, Ls["poses_init"][:n])))
        if not torch.equal(pw, Ls["poses_init"][:n]):
            print(pw.cpu().numpy(), "\n", Ls["poses_init"][:n].cpu().numpy())
        print("dense" if dense else "lowres", "image", i, " | ".join(msg), "status", Ls["hough_status"].tolist(), Lw["hough_status"].tolist())
