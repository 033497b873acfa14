#!/bin/bash
# Run on the GPU box (under gpurun): launch list of the bench command + one full ncu capture
# of the scan kernel (ba_forward2_kernel) on the bench network at one wave of particle-days.
# Outputs go to gpurun_out/; tools/refresh_profiles_r2.sh turns them into profiles/.
CMD="python tools/profile_forward.py 4736 0.012 2 3"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$BENCH > gpurun_out/r2_bench_short.json 2> gpurun_out/r2_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_r2.csv $BENCH > gpurun_out/r2_ncu_launches.log 2>&1
$CMD > gpurun_out/r2_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ba_forward2_kernel -s 2 -c 1 \
    -f -o gpurun_out/prof_r2_forward2 $CMD > gpurun_out/r2_ncu_full.log 2>&1
cat gpurun_out/r2_prof_plain.log
