#!/bin/bash
# 2 GPUs: 8192 x 2048 KH = two 1024-row slabs with neighbours on both sides (periodic y): what a rank of an 8-slab run sees
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --ny 2048 --steps 100 --warmup 10 --reps 3 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=2 ny',d['config']['Ny'],'Mcell/s',round(d['value']),'ms/step',d['ms_per_step'],'ms/launch',d['roofline']['ms_per_launch'],'frac',d['roofline']['frac'],'cflwait',d['roofline']['cfl_mail_wait_us_per_step'], d['config']['ms_per_step_all_repetitions'])"
done
