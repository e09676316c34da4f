mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_network_gpu.py tests/test_pixel_ops_gpu.py -x -q 2>&1 | tail -15
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
cat gpurun_out/bench_full.json; tail -5 gpurun_out/bench_full.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_full.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full_bench.log 2>&1
