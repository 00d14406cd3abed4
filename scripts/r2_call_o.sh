#!/bin/bash
# round-2 GPU call O (2 GPUs): multi mode after making the halo waits convergent
cd "$(dirname "$0")/.."
timeout 300 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -2
run2() {
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29530 + RANDOM % 100)) bench.py --gpus 2 --steps 40 --warmup 5 --e2e-steps 0 --reps 1 --sustained-steps 0 --no-scaling-blocks "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('N=2', '$*', round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'ms/launch', round(r['ms_per_launch'],4), 'frac', round(r['frac'],3), 'cfl wait us/step', round(r['cfl_mail_wait_us_per_step'],2))"
}
run2; run2
run2 --ny 2048; run2 --ny 2048
for ny in 1024 4096; do
BENCH_EXTRA="--ny $ny" scripts/bench_variants.sh main 2>/dev/null | sed 's/^/plain          /'
FV2D_FORCE_MODE=1 FV2D_FORCE_EDGE_RUNS=1 BENCH_EXTRA="--ny $ny" scripts/bench_variants.sh main 2>/dev/null | sed 's/^/multi+edgeruns /'
FV2D_FORCE_MODE=2 BENCH_EXTRA="--ny $ny" scripts/bench_variants.sh main 2>/dev/null | sed 's/^/general        /'
done
