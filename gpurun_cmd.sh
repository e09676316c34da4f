mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'], d['clocks'])
PY
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_conv_row2' -c 3 python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline 2>&1 | grep -E "gpu__time|tensor_cycles|k_conv_row2" | cut -c1-120
