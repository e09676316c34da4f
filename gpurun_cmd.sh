mkdir -p gpurun_out
timeout 900 python tools/bench_ref_kernels.py 2>&1 | tee gpurun_out/bench_ref_kernels.log | tail -8
