mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_ops_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python - <<'PY'
import torch, numpy as np
from posecnn_b200 import train_ops, synth
dev=torch.device('cuda:0')
B,H,W,C=32,480,640,22
sc=synth.make_scene(batch=4,height=H,width=W,num_classes=C,seed=5)
label=torch.from_numpy(np.tile(sc['label'],(8,1,1))).to(dev)
cen=torch.rand((B,C,3),device=dev)*400+1
def ev(fn,n=10):
    fn(); torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
t=ev(lambda: train_ops.generate_vertex_targets(label,cen,10.0)); print(f"vertex_targets B=32 (memset + sparse): {t:.3f} ms ({2*B*H*W*3*C*4/t/1e6:.0f} GB/s)")
PY
