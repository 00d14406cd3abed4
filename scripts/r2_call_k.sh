#!/bin/bash
# round-2 GPU call K (8 GPUs): multi-GPU bitwise tests + the scaling lines at N = 8, 4 (N = 1, 2 are measured on smaller boxes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m pytest tests/test_gpu_multigpu.py -q 2>&1 | tail -4 | tee gpurun_out/r2_pytest_multigpu_8gpu.txt
for n in 8 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/r2_scale_n$n.json 2> gpurun_out/r2_scale_n$n.err
  tail -2 gpurun_out/r2_scale_n$n.err | cut -c1-300
done
python - <<'PY'
import json
for n in (8,4):
    try:
        d=json.loads(open(f'gpurun_out/r2_scale_n{n}.json').read().strip().splitlines()[-1])
        print(n, round(d['value']), 'ms/step', round(d['ms_per_step'],4), 'ms/launch', round(d['roofline']['ms_per_launch'],4), 'frac', round(d['roofline']['frac'],3), 'hash', d['state_hash']['u64'])
        print('   reps', d['config']['ms_per_step_all_repetitions'])
        print('   sustained', d['sustained'] and (round(d['sustained']['value']), round(d['sustained']['frac'],3)))
        print('   strong16384', d['strong_16384'] and (round(d['strong_16384']['value']), round(d['strong_16384']['ms_per_step'],4), round(d['strong_16384']['frac'],3)))
        print('   weak', d['weak'] and (round(d['weak']['value']), round(d['weak']['ms_per_step'],4), round(d['weak']['frac'],3)))
        print('   e2e', d['e2e'] and round(d['e2e']['value']))
    except Exception as e:
        print(n, 'FAILED', e)
PY
