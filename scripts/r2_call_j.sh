#!/bin/bash
# round-2 GPU call J (1 GPU): the whole GPU test suite + the default bench line + the reference arm
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -25 > gpurun_out/r2_pytest_gpu_1gpu.txt; cat gpurun_out/r2_pytest_gpu_1gpu.txt
echo "== default bench"
(time timeout 900 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err) 2>&1 | grep real
echo "== reference arm"
(time timeout 900 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench_default.err) 2>&1 | grep real
cut -c1-600 gpurun_out/r2_bench_reference_arm.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_default.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','steps','gpu_launches','state_hash','clocks'): print(k, d[k])
print('roofline', {k:v for k,v in d['roofline'].items() if k!='fp64'}); print('fp64', d['roofline']['fp64'])
for k in ('sustained','strong_16384','weak','e2e','cpu_baseline'): print(k, d[k])
PY
