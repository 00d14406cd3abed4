#!/bin/bash
# round-2 GPU call S (1 GPU): streamed host path tests + e2e, small-slab breakdown
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -15
echo "== bench (e2e streamed)"
timeout 600 python bench.py --steps 20 --warmup 5 --reps 1 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline > gpurun_out/r2_s_bench.json 2> gpurun_out/r2_s_bench.err
tail -3 gpurun_out/r2_s_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_s_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d["e2e"]))
print("value", d["value"], "frac", d["roofline"]["frac"])
P
echo "== small slab 8192x1024 (one of 8 y-slabs) on one GPU"
for ny in 1024 2048; do
timeout 300 python bench.py --ny $ny --steps 100 --warmup 10 --reps 3 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline --e2e-steps 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ny',d['config']['Ny'],'Mcell/s',round(d['value']),'ms/step',d['ms_per_step'],'ms/launch',d['roofline']['ms_per_launch'],'frac',d['roofline']['frac'])"
done
echo "== timing build, 1024-row slab"
FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 300 python scripts/sweep_timing.py kelvin_helmholtz_8192_plm_hllc 6 1024
FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 300 python scripts/sweep_timing.py kelvin_helmholtz_8192_plm_hllc 6
