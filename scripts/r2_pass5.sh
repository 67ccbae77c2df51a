#!/bin/bash
mkdir -p gpurun_out
scripts/ncu_capture.sh 5 r2e_prof_c5 2 1
scripts/ncu_capture.sh 3 r2e_prof_c3 2
