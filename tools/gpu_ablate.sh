#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ts*.log
HIDVAE_TC_DEBUG=64 timeout 300 python tools/bench_encode.py --tag small --rows 12101 --shape 32,256,3 --encode-only --reps 1 > gpurun_out/ts_small.log 2>> gpurun_out/ablate.err
HIDVAE_TC_DEBUG=64 timeout 300 python tools/bench_encode.py --tag small --rows 12101 --shape 32,256,3 --reps 1 > gpurun_out/ts_small_train.log 2>> gpurun_out/ablate.err
grep -c TS gpurun_out/ts_small.log gpurun_out/ts_small_train.log
