# cost of the pose-refinement choices inside the full C2 step (device-resident throughput; separate processes)
for r in arun lm none; do
  python bench.py --no-cpu --steps 50 --refine $r 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); k=d['kernel_ms_per_step']
print('refine[$r]', d['value'], d['ms_per_step'], {n:v for n,v in k.items() if 'refi' in n or 'enqueue_step' in n})"
done
