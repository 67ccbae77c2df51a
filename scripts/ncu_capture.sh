#!/bin/bash
# usage (on the GPU box, from the repo root): scripts/ncu_capture.sh <config> <tag> [launches] [num_samples|-] [batch]
# Runs the target once without ncu (must exit 0), then captures with --set full + source counters into
# gpurun_out/<tag>.ncu-rep: the render kernel of the LAST launch, or -- batch -- the batched render kernel AND the
# accumulate kernel that follows it (the two kernels of bench.py's timed step).
cfg=$1; tag=$2; launches=${3:-2}; ns=${4:--}; batch=${5:-}
python scripts/profile_target.py $cfg $launches $ns $batch > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; cat gpurun_out/${tag}_plain.log; exit 1; }
cat gpurun_out/${tag}_plain.log
if [ "$batch" = "batch" ]; then
  ncu --set full --import-source on --clock-control none -k regex:'render_kernel|accumulate_kernel' -s 2 -c 2 -f \
      -o gpurun_out/${tag} python scripts/profile_target.py $cfg $launches $ns batch > gpurun_out/${tag}_ncu.log 2>&1
else
  ncu --set full --import-source on --clock-control none -k regex:render_kernel -s $((launches-1)) -c 1 -f \
      -o gpurun_out/${tag} python scripts/profile_target.py $cfg $launches $ns > gpurun_out/${tag}_ncu.log 2>&1
fi
tail -2 gpurun_out/${tag}_ncu.log
