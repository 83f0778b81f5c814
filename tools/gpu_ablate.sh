#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ts*.log
HIDVAE_TC_DEBUG=64 timeout 300 python tools/bench_encode.py --tag big --rows 4194304 --shape 32,256,3 --encode-only --reps 1 > gpurun_out/ts_big.log 2>> gpurun_out/ablate.err
grep -c TS gpurun_out/ts_big.log
