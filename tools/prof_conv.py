"""One launch of each conv tile shape for ncu (--set full): conv1_2 (BN 64), conv2_2 (BN 128), conv3_2 (BN 256), conv1_1 fused."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from posecnn_b200 import conv
dev = torch.device("cuda:0")
B = 32
for name, H, W, Cin, Cout in [("conv1_2", 480, 640, 64, 64), ("conv2_2", 240, 320, 128, 128), ("conv3_2", 120, 160, 256, 256)]:
    x = torch.randn((B, H, W, Cin), device=dev).to(torch.bfloat16)
    w = conv.hwio_to_tc(torch.randn((3, 3, Cin, Cout), device=dev) * 0.05)
    b = torch.zeros((Cout,), device=dev)
    for _ in range(2):
        y = conv.conv_bf16(x, w, b, 3, True)
    torch.cuda.synchronize()
img = torch.randint(0, 256, (B, 480, 640, 3), dtype=torch.uint8, device=dev)
w1 = conv.conv1_1_weights_to_tc(torch.randn((3, 3, 3, 64), device=dev) * 0.1)
for _ in range(2):
    y = conv.conv1_fused(img, w1, torch.zeros(64, device=dev), (102.9801, 115.9465, 122.7717))
torch.cuda.synchronize()
