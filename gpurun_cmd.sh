mkdir -p gpurun_out
for s in 20 100; do timeout 300 python bench.py --workload hough --batch 1 --steps $s --warmup 10 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('steps', d['steps'], 'ms', d['ms_per_step'])"; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 python bench.py --workload hough --batch 1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep -E "^  [a-zA-Z].*\(|gpu__time" | paste - - | awk '{print $NF, $1, $2}' | tail -18
