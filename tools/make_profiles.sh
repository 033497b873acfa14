#!/bin/bash
# Run on the GPU box (under gpurun): launch list + one full ncu capture of the forward
# kernel for the bench workload at a reduced particle count.  Outputs go to gpurun_out/.
set -e
CMD="python tools/profile_forward.py 2960 0.012 2 3"
$CMD > gpurun_out/profile_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_r1.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/profile_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ba_forward_kernel -s 2 -c 1 \
    -o gpurun_out/prof_r1_forward $CMD > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/profile_plain.log
