"""conv1_1 (k_conv1_tc) alone: device time and write bandwidth (events, L2 flushed).  python tools/prof_conv1.py [batch]
Round-2 experiment recorded in DESIGN.md §4: an epilogue that copied the staged tile with coalesced st.global.cs stores instead of
one TMA tensor store per tile ran at 0.469 ms against 0.318 ms at batch 32 and was removed again."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posecnn_b200 import conv
dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
img = torch.randint(0, 256, (B, 480, 640, 3), dtype=torch.uint8, device=dev)
w1 = conv.conv1_1_weights_to_tc(torch.randn((3, 3, 3, 64), device=dev) * 0.1)
b = torch.zeros(64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    y = conv.conv1_fused(img, w1, b, (102.9801, 115.9465, 122.7717))
ts = []
for _ in range(10):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); y = conv.conv1_fused(img, w1, b, (102.9801, 115.9465, 122.7717)); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print("conv1_1 batch %d: %.4f ms, %.2f TB/s of writes" % (B, ms, y.numel() * 2 / ms / 1e9))
