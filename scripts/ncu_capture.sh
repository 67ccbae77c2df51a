#!/bin/bash
# usage (on the GPU box, from the repo root): scripts/ncu_capture.sh <config> <tag> [launches] [num_samples]
# Runs the target once without ncu (must exit 0), then captures the render kernel of the LAST launch with
# --set full + source counters into gpurun_out/<tag>.ncu-rep.
cfg=$1; tag=$2; launches=${3:-2}; ns=${4:-}
python scripts/profile_target.py $cfg $launches $ns > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; cat gpurun_out/${tag}_plain.log; exit 1; }
cat gpurun_out/${tag}_plain.log
ncu --set full --import-source on --clock-control none -k regex:render_kernel -s $((launches-1)) -c 1 -f \
    -o gpurun_out/${tag} python scripts/profile_target.py $cfg $launches $ns > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
