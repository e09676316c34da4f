import ctypes, os, subprocess
import torch
here = os.path.dirname(os.path.abspath(__file__))
so = os.path.join(here, "libumma_probe_mn.so")
src = os.path.join(here, "umma_probe_mn.cu")
if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-o", so, src])
lib = ctypes.CDLL(so)
torch.manual_seed(0)
A = torch.randn(80, 128, device="cuda").to(torch.bfloat16)      # [pixel][Cout]
B = torch.randn(80, 64, device="cuda").to(torch.bfloat16)       # [pixel][Cin]
out = torch.zeros(2, 9, 128, 64, device="cuda")
rc = lib.umma_probe_mn(ctypes.c_void_p(A.data_ptr()), ctypes.c_void_p(B.data_ptr()), ctypes.c_void_p(out.data_ptr()))
print("rc", rc)
for var in range(1):
    res = []
    for s in range(9):
        want = A[:64].float().t() @ B[s:s + 64].float()            # D[m][n] = sum_p A[p][m] B[p + s][n]
        err = (out[var, s] - want).abs().max().item()
        res.append(f"s{s}:{'OK' if err < 1e-2 else f'{err:.1f}'}")
    print("variant", var, "(LBO=block, SBO=1024)" if var == 0 else "(LBO=1024, SBO=block)", " ".join(res))
