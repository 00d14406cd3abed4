#!/bin/bash
# round-2 GPU call A (1 GPU): full-size parity against the reference binary, ring-cut experiment, sustained baseline
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2; df -h /tmp | tail -1
python -c "import h5py; print('h5py', h5py.__version__)" 2>&1 | tail -1
timeout 900 python scripts/parity_fullsize.py --out gpurun_out/r2_parity_fullsize_v6.json 2>&1 | grep -v WARNING | cut -c1-1500
BENCH_EXTRA="" scripts/bench_variants.sh main cut2 cut4
timeout 300 python bench.py --steps 200 --warmup 10 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks 2>&1 | tail -1 > gpurun_out/r2_v6_sustained200.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_v6_sustained200.json'))
print('sustained200', d['value'], d['ms_per_step'], d['roofline']['frac'], d['clocks'])
PY
timeout 600 python -m pytest tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
