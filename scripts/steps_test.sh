cd /root/repo
for s in 20 40 100 200 400; do
python bench.py --steps $s --warmup 5 --e2e-steps 0 --no-cpu-baseline --reps 1 --sustained-steps 0 --no-scaling-blocks 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('steps',d['steps'],'ms/step %.4f'%d['ms_per_step'],'ms/launch %.4f'%d['roofline']['ms_per_launch'],d['clocks'])"
done
nvidia-smi -q -d POWER | grep -E "Power Limit|Power Draw|Default" | head -8
nvidia-smi -q -d CLOCK | head -30
