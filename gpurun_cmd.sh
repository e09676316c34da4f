mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_network_gpu.py tests/test_pixel_ops_gpu.py -m gpu -x -q 2>&1 | tail -8
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'], d['clocks'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_up8_heads|k_lowres_heads|k_roi_pool' -c 8 python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline 2>&1 | grep -E "k_up8|k_lowres|k_roi_pool|gpu__time" | head -20
