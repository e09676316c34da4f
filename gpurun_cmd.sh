mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_conv_gpu.py -m gpu -x -q -k conv1 2>&1 | tail -3
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_conv1_tc' -c 1 python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline 2>&1 | grep -E "gpu__time|inst_executed|issue_active"
