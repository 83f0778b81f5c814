#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ts16*.log
HIDVAE_TC_DEBUG=64 timeout 300 python tools/bench_encode.py --tag big --rows 4194304 --shape 32,256,3 --encode-only --reps 1 > gpurun_out/ts16_big.log 2>> gpurun_out/ts16.err
grep -c TS gpurun_out/ts16_big.log
timeout 300 python tools/bench_encode.py --tag big --rows 4194304 --shape 32,256,3 --encode-only --reps 10 > gpurun_out/ts16_time.log 2>> gpurun_out/ts16.err
cat gpurun_out/ts16_time.log
