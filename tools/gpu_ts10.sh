#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ts10*.log
export HIDVAE_B200_LIB=$PWD/hid-vae_b200/build/variants/${TSLIB:-s3i}.so
HIDVAE_TC_DEBUG=${TSDBG:-64} timeout 300 python tools/bench_encode.py --tag big --rows 4194304 --shape 32,256,3 --encode-only --reps 1 > gpurun_out/ts10_big.log 2>> gpurun_out/ts10.err
HIDVAE_TC_DEBUG=${TSDBG:-64} timeout 300 python tools/bench_encode.py --tag small --rows 12101 --shape 32,256,3 --encode-only --reps 1 > gpurun_out/ts10_small.log 2>> gpurun_out/ts10.err
grep -c TS gpurun_out/ts10_big.log gpurun_out/ts10_small.log
