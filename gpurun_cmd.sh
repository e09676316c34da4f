mkdir -p gpurun_out
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_nms_gpu.py tests/test_pixel_ops_gpu.py tests/test_hough_gpu.py -m gpu -x -q -k "not batch32 and not full_size" 2>&1 | tail -6 ) > gpurun_out/sanitizer_ops.txt; cat gpurun_out/sanitizer_ops.txt
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_network_gpu.py tests/test_conv_gpu.py -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/sanitizer_net.txt; cat gpurun_out/sanitizer_net.txt
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_nms_gpu.py -m gpu -x -q -k "golden or ties" 2>&1 | tail -5 ) > gpurun_out/sanitizer_race_nms.txt; cat gpurun_out/sanitizer_race_nms.txt
