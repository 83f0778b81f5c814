#!/bin/bash
# One ncu --set full capture of the large fused-forward launches (encode + train forward) of the current build.
mkdir -p gpurun_out; rm -f gpurun_out/status.txt
timeout 300 python tools/profile_step.py --rows 1048576 --reps 2 > gpurun_out/prof_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:rq_fwd_tc" -s 4 -c 3 -o gpurun_out/prof_tc -f python tools/profile_step.py --rows 1048576 --reps 2 > gpurun_out/ncu_full.log 2>&1
echo "ncu_full rc=$?" >> gpurun_out/status.txt
cat gpurun_out/status.txt; tail -3 gpurun_out/ncu_full.log
