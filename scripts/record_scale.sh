#!/bin/bash
# Developer tool (GPU box with N GPUs): multi-GPU bitwise tests (N = 8 only) + the driver-like bench line at N
#   scripts/record_scale.sh N [tag]
N=${1:-8}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
T=${2:-r2v9}
if [ "$N" == "8" ]; then
  timeout 600 python -m pytest tests/test_gpu_multigpu.py -q 2>&1 | tail -3 | tee gpurun_out/${T}_pytest_multigpu_8gpu.txt
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/${T}_scale_n$N.err | tail -1 > gpurun_out/${T}_scale_n$N.json
tail -3 gpurun_out/${T}_scale_n$N.err
python - gpurun_out/${T}_scale_n$N.json <<'PY'
import json,sys
d=json.load(open(sys.argv[1]))
print("N",d["n_gpus"],"value",round(d["value"]),"ms/step",d["ms_per_step"],"frac",d["roofline"]["frac"],"hash",d["state_hash"]["u64"],"cflwait",d["roofline"]["cfl_mail_wait_us_per_step"])
for k in ("sustained","strong_16384","weak","e2e"):
    b=d.get(k) or {}
    print("  ",k,round(b.get("value",0)),b.get("ms_per_step"),b.get("frac"))
PY
