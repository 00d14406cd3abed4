#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== PDL A/B"
for ny in 1024 8192; do for nopdl in 0 1; do
if [ $nopdl = 1 ]; then export FV2D_NO_PDL=1; else unset FV2D_NO_PDL; fi
timeout 300 python bench.py --ny $ny --steps 100 --warmup 10 --reps 3 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('nopdl=$nopdl ny',d['config']['Ny'],'Mcell/s',round(d['value']),'ms/step',round(d['ms_per_step'],5),'frac',round(d['roofline']['frac'],4),'bracketed',d['roofline']['event_bracketed_launches'], d['state_hash']['u64'])"
done; done
unset FV2D_NO_PDL
echo "== e2e trace"
FV2D_STREAM_TRACE=1 timeout 300 python bench.py --steps 5 --warmup 3 --reps 1 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline --e2e-steps 4 2>&1 | grep -v WARNING | cut -c1-400 | tail -9
