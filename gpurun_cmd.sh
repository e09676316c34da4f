mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_network_gpu.py -x -q 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python -c "import json;d=json.load(open('gpurun_out/bench_full.json'));print(d['value'],d['ms_per_step'],d['breakdown_ms'],d['roofline']['frac'])"; tail -5 gpurun_out/bench_full.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-graph > gpurun_out/ncu_full_bench.log 2>&1
