#!/bin/bash
# round-2 GPU call H (2 GPUs): multi-GPU correctness of the folded halo exchange + the new bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -5
echo "== bench N=1"
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-steps 2 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
echo "== bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -2 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads(open(f'gpurun_out/r2_bench_n{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value']), d['ms_per_step'], 'frac', round(d['roofline']['frac'],3), 'hash', d['state_hash'], 'sustained', d['sustained'] and (round(d['sustained']['value']), round(d['sustained']['frac'],3)), 'strong16384', d['strong_16384'] and (round(d['strong_16384']['value']), round(d['strong_16384']['frac'],3)), 'weak', d['weak'] and (round(d['weak']['value']), round(d['weak']['frac'],3)), 'e2e', d['e2e'] and round(d['e2e']['value']), 'launches', d['gpu_launches'], 'fp64', d['roofline'].get('fp64'), 'cpu', d['cpu_baseline'])
    except Exception as e:
        print(n, 'FAILED', e)
PY
