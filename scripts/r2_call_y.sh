#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== schedule knobs on a 1024-row slab (kPlain, edge runs forced as on a slab with neighbours)"
export FV2D_FORCE_EDGE_RUNS=1
for knobs in "220 96" "150 96" "180 96" "300 96" "150 128" "120 128" "100 128"; do
set -- $knobs
FV2D_SCHED_C100=$1 FV2D_SCHED_HMAX=$2 timeout 300 python bench.py --ny 1024 --steps 100 --warmup 10 --reps 5 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('C100=$1 HMAX=$2 ny',d['config']['Ny'],'Mcell/s',round(d['value']),'ms/step',round(d['ms_per_step'],5),'all',[round(x,5) for x in d['config']['ms_per_step_all_repetitions']])"
done
