#!/bin/bash
# Developer tool (GPU box): sensitivity of the sweep to the chunk height (FV2D_CHUNK_ROWS).
cd "$(dirname "$0")/.."
for cr in "$@"; do
  if [ "$cr" == "auto" ]; then unset FV2D_CHUNK_ROWS; else export FV2D_CHUNK_ROWS=$cr; fi
  python bench.py --steps 40 --warmup 5 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('chunk_rows=$cr', round(d['value']), 'ms/launch %.4f'%d['roofline']['ms_per_launch'], 'frac %.4f'%d['roofline']['frac'])"
done
