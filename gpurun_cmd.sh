mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python -c "import json;d=json.load(open('gpurun_out/bench_full.json'));print(d['value'],d['ms_per_step'],d['breakdown_ms'],d['roofline']['frac'],d['e2e'],d['clocks'])"; tail -5 gpurun_out/bench_full.err
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-graph --no-e2e > gpurun_out/bench_full_eager.json 2>> gpurun_out/bench_full.err
python -c "import json;d=json.load(open('gpurun_out/bench_full_eager.json'));print('eager',d['value'],d['ms_per_step'])"
