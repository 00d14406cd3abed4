#!/bin/bash
# Developer tool (GPU box with N GPUs): end-to-end rate of N ranks with and without NUMA-local host pages
N=${1:-2}
cd "$(dirname "$0")/.."
for nobind in 0 1; do
if [ $nobind = 1 ]; then export FV2D_NO_NUMA_BIND=1; else unset FV2D_NO_NUMA_BIND; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus $N --steps 5 --warmup 3 --reps 1 --sustained-steps 0 --no-scaling-blocks --e2e-steps 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nobind=$nobind N',d['n_gpus'],'e2e',round(d['e2e']['value']),'numa',d['config']['host_numa_binding'])"
done
nvidia-smi topo -m 2>/dev/null | head -14
lscpu | grep -i -E "numa|socket|model name" | head -8
