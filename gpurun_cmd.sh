mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/full_launches.csv python bench.py --steps 1 --warmup 1 --no-graph --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; tail -1 gpurun_out/ncu_full.log | cut -c1-100
python profiles/step_breakdown.py gpurun_out/full_launches.csv profiles/r01_full_b32 | tail -5
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 1500 gpurun_out/bench_full.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 900 gpurun_out/bench_reference.json
cp profiles/r01_full_b32_step_breakdown.txt profiles/r01_full_b32_trunk_traffic.json gpurun_out/
