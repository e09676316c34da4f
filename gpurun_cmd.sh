mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_conv_gpu.py -x -q 2>&1 | tail -30 > gpurun_out/pytest_conv.log
tail -30 gpurun_out/pytest_conv.log
