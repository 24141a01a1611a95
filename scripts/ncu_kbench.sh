#!/bin/bash
# usage: scripts/ncu_kbench.sh <tag> <kbench-arg> <kernel-regex>   (run under gpurun)
set -u
tag=$1; what=$2; k=$3
mkdir -p gpurun_out
python scripts/kbench.py $what > gpurun_out/${tag}_plain.log 2>&1 || { tail -5 gpurun_out/${tag}_plain.log; exit 1; }
cat gpurun_out/${tag}_plain.log
ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -o gpurun_out/${tag} -f python scripts/kbench.py $what > gpurun_out/${tag}_ncu.log 2>&1
ls -la gpurun_out/${tag}.ncu-rep
