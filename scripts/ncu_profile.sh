#!/bin/bash
# Run on the GPU box (under gpurun): launch list + full captures of the three heaviest kernels of one bench step.
# Usage: scripts/ncu_profile.sh <tag>     -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{hamming,score,remap}.ncu-rep
set -u
tag=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu"
KREGEX='regex:remap_kernel|hamming_|match_select|stereo_lift|score_kernel|hypothesize|argmax_kernel|mask_kernel|refit_kernel|gather_desc|assemble_kernel|segments_kernel|stats_kernel'
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 500 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
for k in hamming_partial score_kernel remap_kernel; do
  short=${k%%_*}
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 2 -o gpurun_out/${tag}_${short} -f $CMD > gpurun_out/${tag}_ncu_${short}.log 2>&1
done
ls -la gpurun_out/
