#!/bin/bash
# Run on the GPU box (under gpurun).  ONE ncu pass per call, and only after the same command exited 0 without ncu.
# Usage: scripts/ncu_profile.sh <tag> launches            -> gpurun_out/<tag>_launches.csv   (per-launch durations of one step)
#        scripts/ncu_profile.sh <tag> <kernel-regex> <short> -> gpurun_out/<tag>_<short>.ncu-rep (--set full, 2 launches)
set -u
tag=${1:-r01}
what=${2:-launches}
short=${3:-$what}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras ${SOS_BENCH_ARGS:-}"
KREGEX='regex:remap|hamming_|mma_kernel|expand_kernel|match_select|stereo_|score_kernel|hypothesize|argmax_kernel|mask_kernel|refit_kernel|refine_kernel|gather_desc|assemble_kernel|segments_kernel|stats_kernel|carry_over'
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/${tag}_plain.log; exit 1; }
tail -c 600 gpurun_out/${tag}_plain.log; echo
if [ "$what" = launches ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 600 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_launches.log 2>&1
  wc -l gpurun_out/${tag}_launches.csv
else
  ncu --set full --clock-control none --import-source on -k regex:$what -s 6 -c 2 -o gpurun_out/${tag}_${short} -f $CMD > gpurun_out/${tag}_ncu_${short}.log 2>&1
  ls -la gpurun_out/${tag}_${short}.ncu-rep
fi
