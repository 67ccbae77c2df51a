#!/bin/bash
mkdir -p gpurun_out
scripts/ncu_capture.sh 3 r2g_prof_c3 2
