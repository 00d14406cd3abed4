#!/bin/bash
# round-2 GPU call T (1 GPU): streamed host path tests, host-link ceiling, small-slab timing breakdown
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream.py -x -q 2>&1 | tail -5
echo "== host link"
timeout 300 python scripts/pcie_probe.py 2>&1 | tee gpurun_out/r2_pcie_probe.txt
echo "== e2e with other block heights"
for rows in 128 512 1024; do
FV2D_STREAM_ROWS=$rows timeout 300 python bench.py --steps 5 --warmup 3 --reps 1 --sustained-steps 0 --no-scaling-blocks --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('rows $rows e2e',round(d['e2e']['value']),'ms',d['e2e']['ms_per_step'],'serial',round(d['e2e']['serial']['value']))"
done
echo "== timing build, 1024-row slab"
FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 300 python scripts/sweep_timing.py kelvin_helmholtz_8192_plm_hllc 6 1024 2>&1 | grep -v WARNING
FV2D_B200_LIB=$PWD/scratch/lib_timing.so timeout 300 python scripts/sweep_timing.py kelvin_helmholtz_8192_plm_hllc 6 2>&1 | grep -v WARNING
