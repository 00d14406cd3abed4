#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -5
for rows in 256 512; do
FV2D_STREAM_ROWS=$rows timeout 300 python bench.py --steps 5 --warmup 3 --reps 1 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rows $rows e2e',round(d['e2e']['value']),'ms',d['e2e']['ms_per_step'],'streamed',d['e2e']['steps_streamed'],'serial',round(d['e2e']['serial']['value']))"
done
