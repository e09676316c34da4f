"""Weight-gradient kernel (k_wgrad_tc) next to the forward kernel on the same layer shapes: device time (events, L2 flushed)
and TFLOP/s.   python tools/prof_wgrad.py [batch] [--few]   (--few: two layers, one call each, for `ncu --set full`)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posecnn_b200 import backward, conv
B = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 16
few = "--few" in sys.argv
dev = torch.device("cuda:0")
layers = [("conv1_2", 480, 640, 64, 64), ("conv2_1", 240, 320, 64, 128), ("conv2_2", 240, 320, 128, 128), ("conv3_1", 120, 160, 128, 256),
          ("conv3_2", 120, 160, 256, 256), ("conv4_1", 60, 80, 256, 512), ("conv4_2", 60, 80, 512, 512), ("conv5_1", 30, 40, 512, 512)]
if few:
    layers = [l for l in layers if l[0] in ("conv1_2", "conv4_2")]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=5):
    for _ in range(1 if few else 2):
        fn()
    if few:
        torch.cuda.synchronize()
        return 0.0
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


for name, H, W, Cin, Cout in layers:
    x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
    dz = (torch.randn((B, H, W, Cout), device=dev) * 0.1).to(torch.bfloat16)
    w = conv.hwio_to_tc(torch.randn((3, 3, Cin, Cout), device=dev) * 0.05)
    b = torch.zeros((Cout,), device=dev)
    out = torch.empty((B, H, W, Cout), dtype=torch.bfloat16, device=dev)
    dW = torch.empty((Cout, 9 * Cin), dtype=torch.float32, device=dev)
    fl = 2.0 * B * H * W * 9 * Cin * Cout
    t_f = timed(lambda: conv.conv_bf16(x, w, b, 3, True, 0, out))
    t_w = timed(lambda: backward.conv_wgrad(x, dz, 3, out=dW))
    if not few:
        print(f"{name:8s} B={B} {H}x{W} {Cin}->{Cout}: forward {t_f:.3f} ms {fl / t_f / 1e9:6.0f} TF/s | wgrad (+finish) {t_w:.3f} ms {fl / t_w / 1e9:6.0f} TF/s"
              f" | ratio {t_w / t_f:.2f}", flush=True)
