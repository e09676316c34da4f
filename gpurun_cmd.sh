mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_ops_gpu.py -m gpu -x -q 2>&1 | tail -6
timeout 300 python - <<'PY'
import torch, numpy as np
from posecnn_b200 import train_ops, synth
dev=torch.device('cuda:0')
B,H,W,C=32,480,640,22
sc=synth.make_scene(batch=4,height=H,width=W,num_classes=C,seed=5)
label=torch.from_numpy(np.tile(sc['label'],(8,1,1))).to(dev)
cen=torch.rand((B,C,3),device=dev)*400+1
pred=torch.randn((B,H,W,3*C),device=dev)
def ev(fn,n=10):
    fn(); torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
def unfused():
    t,w=train_ops.generate_vertex_targets(label,cen,10.0); return train_ops.smooth_l1_loss_vertex(pred,t,w,1.0)
print(f"vertex loss B=32: targets + smooth L1 {ev(unfused):.3f} ms   fused {ev(lambda: train_ops.vertex_loss_from_centers(pred,label,cen,10.0,1.0)):.3f} ms   fused + grad {ev(lambda: train_ops.vertex_loss_from_centers(pred,label,cen,10.0,1.0,want_grad=True)):.3f} ms")
PY
for s in 20 100; do timeout 300 python bench.py --workload hough --batch 1 --steps $s --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hough b1 steps', d['steps'], 'ms', d['ms_per_step'], d['config']['cuda_graph'], d['clocks']['samples'])"; done
timeout 300 python bench.py --workload hough --steps 50 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_hough.json; python -c "
import json; d=json.loads(open('gpurun_out/bench_hough.json').read().strip().splitlines()[-1]); print('hough b32', d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])"
timeout 300 python bench.py --workload hough --batch 1 --steps 100 --warmup 10 --no-cpu-baseline 2>/dev/null > gpurun_out/bench_hough_b1.json
