#!/bin/bash
mkdir -p gpurun_out
scripts/ncu_capture.sh 2 r2j_prof_c2 16 - batch
