#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
for lib in prev new prev new; do
  if [ $lib = prev ]; then export FV2D_B200_LIB=$PWD/scratch/lib_prev.so; else unset FV2D_B200_LIB; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --ny 2048 --steps 100 --warmup 10 --reps 3 --rep-pause 0.5 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$lib N=2 ny=2048',round(d['value']),'ms/step',[round(x,5) for x in d['config']['ms_per_step_all_repetitions']],'cflwait',round(d['roofline']['cfl_mail_wait_us_per_step'],2), d['state_hash']['u64'])"
done
