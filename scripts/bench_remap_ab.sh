for v in "" "SOS_REMAP_V1=1"; do
  env $v python bench.py --no-cpu --steps 50 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('$v', d['value'], d['kernels']['remap']['ms'], d['kernels']['remap']['frac'])"
done
