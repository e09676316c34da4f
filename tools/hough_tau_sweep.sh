#!/bin/bash
# Hough k_vote: sweep of the end-point re-check margin tau0 (exactness tests + timing for each value)
for t in 0.03 0.008 0.004 0.002; do
  echo "=== PCNN_HOUGH_TAU0=$t"
  PCNN_HOUGH_TAU0=$t timeout 900 python -m pytest tests/test_hough_gpu.py tests/test_reference_kernels_gpu.py -m gpu -q 2>&1 | tail -2
  PCNN_HOUGH_TAU0=$t timeout 300 python bench.py --workload hough --steps 20 --warmup 3 --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('hough b32 ms', d['ms_per_step'], 'frac', d['roofline']['frac'])"
done
