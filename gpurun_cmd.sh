mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pixel_ops_gpu.py tests/test_network_gpu.py tests/test_nms_gpu.py -m gpu -x -q 2>&1 | tail -4
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'], d['clocks'], d['roofline']['traffic'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_roi_pool' -c 2 python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline 2>&1 | grep -E "gpu__time"
