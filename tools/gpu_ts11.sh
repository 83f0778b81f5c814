#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ts11*.log
export HIDVAE_B200_LIB=$PWD/hid-vae_b200/build/variants/${TSLIB:-v11i}.so HIDVAE_TC_IMPL=v11
HIDVAE_TC_DEBUG=64 timeout 300 python tools/bench_encode.py --tag big --rows 4194304 --shape 32,256,3 --encode-only --reps 1 > gpurun_out/ts11_big.log 2>> gpurun_out/ts11.err
HIDVAE_TC_DEBUG=64 timeout 300 python tools/bench_encode.py --tag small --rows 12101 --shape 32,256,3 --encode-only --reps 1 > gpurun_out/ts11_small.log 2>> gpurun_out/ts11.err
grep -c TS gpurun_out/ts11_big.log gpurun_out/ts11_small.log
