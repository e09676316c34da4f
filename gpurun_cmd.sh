mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_network_gpu.py tests/test_conv_gpu.py -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_net.log
tail -40 gpurun_out/pytest_net.log
