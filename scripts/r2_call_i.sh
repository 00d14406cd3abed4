#!/bin/bash
# round-2 GPU call I (2 GPUs): multi-GPU variant after hoisting the halo waits out of the row loop
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
for rep in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$rep bench.py --gpus 2 --steps 40 --warmup 5 --e2e-steps 0 --reps 1 --sustained-steps 0 --no-scaling-blocks 2>/dev/null | tail -1 > gpurun_out/r2_n2_$rep.json
python - $rep <<'PY'
import json,sys
d=json.loads(open(f'gpurun_out/r2_n2_{sys.argv[1]}.json').read())
print('N=2', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'ms/launch', round(d['roofline']['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],3))
PY
done
scripts/bench_variants.sh main 2>/dev/null
BENCH_EXTRA="--ny 4096" scripts/bench_variants.sh main 2>/dev/null
