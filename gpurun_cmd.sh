mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo rc=$?; wc -l gpurun_out/bench_2gpu.json; python - <<'PY'
import json
for l in open('gpurun_out/bench_2gpu.json'):
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['config'].get('parallelism'))
    else: print('NON-JSON LINE:', l[:200])
PY
tail -3 gpurun_out/bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -c 600
