set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_hough.json 2> gpurun_out/bench_hough.err
timeout 300 python bench.py --steps 20 --warmup 3 --batch 1 --no-e2e --no-cpu-baseline > gpurun_out/bench_hough_b1.json 2>> gpurun_out/bench_hough.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_hough.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_vote" -s 3 -c 1 -o gpurun_out/prof_hough python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_hough.json gpurun_out/bench_hough_b1.json
