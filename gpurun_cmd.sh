mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pixel_ops_gpu.py tests/test_reference_kernels_gpu.py -x -q 2>&1 | tail -6
timeout 600 python tools/bench_ops.py 2>&1 | tee gpurun_out/bench_ops.log | tail -4
