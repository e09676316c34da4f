mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_network_gpu.py -m gpu -x -q 2>&1 | tail -15
