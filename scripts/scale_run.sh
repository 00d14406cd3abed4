#!/bin/bash
# Developer tool (multi-GPU box): N-GPU bitwise tests + strong-scaling bench lines -> gpurun_out/scale_*.json
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multigpu.py -x -q 2>&1 | tail -3
run() { # workload N steps
  local wl=$1 n=$2 steps=$3 port=$((29500 + RANDOM % 400))
  if [ "$n" == "1" ]; then
    timeout 300 python bench.py --workload $wl --steps $steps --warmup 5 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks 2>&1 | tail -1 > gpurun_out/scale_${wl}_n$n.json
  else
    timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port \
      bench.py --gpus $n --workload $wl --steps $steps --warmup 5 2>&1 | tail -1 > gpurun_out/scale_${wl}_n$n.json
  fi
  python - gpurun_out/scale_${wl}_n$n.json <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(f"{d['config']['workload']:36s} N={d['n_gpus']} {d['value']:9.0f} Mcell/s  ms/step={d['ms_per_step']:.4f} sweep_frac={d['roofline']['frac']:.3f} share={d['roofline']['share_of_step']:.3f}")
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1]).read()[-600:])
PY
}
for spec in "$@"; do
  IFS=: read wl n steps <<< "$spec"
  run $wl $n $steps
done
