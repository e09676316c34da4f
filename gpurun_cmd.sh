mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; python - <<'PY'
import json
ls=open('gpurun_out/bench_full.json').read().strip().splitlines(); print('lines', len(ls))
d=json.loads(ls[-1]); print(d['value'], d['ms_per_step'], d['steps'], d['breakdown_ms'], d['e2e']['value'], d['clocks'], d['roofline']['frac'], d['roofline']['traffic'], d['gpu_launches']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
