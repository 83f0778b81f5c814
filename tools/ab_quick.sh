#!/bin/bash
# tools/ab_quick.sh WHAT variant ...: bench_kernels.py WHAT of the shipped build and of the named A/B builds (no tests)
cd "$(dirname "$0")/.."
what=$1; shift
: > gpurun_out/ab_quick.jsonl
timeout 200 python tools/bench_kernels.py $what --tag shipped >> gpurun_out/ab_quick.jsonl 2>gpurun_out/ab_quick.err
for v in "$@"; do
  HIDVAE_B200_LIB=$PWD/hid-vae_b200/build/variants/$v.so timeout 200 python tools/bench_kernels.py $what --tag $v >> gpurun_out/ab_quick.jsonl 2>>gpurun_out/ab_quick.err
done
cat gpurun_out/ab_quick.jsonl; tail -3 gpurun_out/ab_quick.err
