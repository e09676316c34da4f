mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_hough_gpu.py tests/test_reference_kernels_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --workload hough --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_hough.json 2> gpurun_out/bench_hough.err; python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_hough.json').read().strip().splitlines()[-1]); print('hough', d['value'], d['ms_per_step'], d['roofline']['frac'])
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python - <<'PY'
import json; d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['breakdown_ms'], d['e2e']['value'])
PY
