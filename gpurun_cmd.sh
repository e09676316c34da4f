mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_network_gpu.py -x -q 2>&1 | tail -6
timeout 300 python tools/bench_conv.py 32 2>&1 | tee gpurun_out/bench_conv.log | head -4
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
python -c "import json;d=json.load(open('gpurun_out/bench_full.json'));print(d['value'],d['ms_per_step'],d['breakdown_ms'],d['roofline']['frac'])"; tail -5 gpurun_out/bench_full.err
