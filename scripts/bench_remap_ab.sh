# A/B of the three 3-channel remap kernels inside the full C2 step (separate processes: the switch is read once)
for v in "" "SOS_REMAP_V2=1" "SOS_REMAP_V1=1"; do
  env $v python bench.py --no-cpu --steps 50 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('variant[$v]', d['value'], d['kernels']['remap']['ms'], d['kernels']['remap']['frac'])"
done
