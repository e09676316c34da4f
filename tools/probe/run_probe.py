import ctypes, os, subprocess, sys
import torch
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "libumma_probe.so")
if not os.path.exists(so):
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", so, os.path.join(here, "umma_probe.cu")])
lib = ctypes.CDLL(so)
torch.manual_seed(0)
rowsA = 136
A = torch.randn(rowsA, 64, device="cuda").to(torch.bfloat16)
B = torch.randn(64, 64, device="cuda").to(torch.bfloat16)
out = torch.zeros(8, 8, 128, 64, device="cuda")
rc = lib.umma_probe(ctypes.c_void_p(A.data_ptr()), rowsA, ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(out.data_ptr()))
print("rc", rc)
for s in range(8):
    want = A[s:s + 128].float() @ B.float().t()
    res = []
    for bo in range(8):
        err = (out[s, bo] - want).abs().max().item()
        res.append(f"bo{bo}:{'OK' if err < 1e-2 else f'{err:.1f}'}")
    print(f"shift {s} rows:", " ".join(res))
