# A/B of the 3-channel remap kernels inside the full C2 step (separate processes: the switches are read once)
#   default = per-frame 4x8 patches; TMA = batch-looped blocks with TMA-staged LUT tiles (+ PROBE 1: no gathers, 2: no
#   stores, 3: neither); V2 = per-frame row mapping; V1 = generic 4 px/thread
for v in "" "SOS_REMAP_TMA=1" "SOS_REMAP_TMA=1 SOS_REMAP_PROBE=1" "SOS_REMAP_TMA=1 SOS_REMAP_PROBE=2" "SOS_REMAP_TMA=1 SOS_REMAP_PROBE=3" "SOS_REMAP_V2=1" "SOS_REMAP_V1=1"; do
  env $v python bench.py --no-cpu --steps 50 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1]); print('variant[$v]', d['value'], d['kernels']['remap']['ms'], d['kernels']['remap']['frac'])"
done
