#!/bin/bash
# A/B timing of the variant libraries under hid-vae_b200/build/variants (encode, 4 Mi and 12,101 rows)
mkdir -p gpurun_out; rm -f gpurun_out/ab.jsonl gpurun_out/ab.err
for lib in hid-vae_b200/build/variants/*.so; do
  name=$(basename $lib .so)
  for rows in 4194304 12101; do
    HIDVAE_B200_LIB=$PWD/$lib timeout 300 python tools/bench_encode.py --tag $name --rows $rows --shape 32,256,3 --encode-only --reps 20 >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
  done
  case $name in *i) HIDVAE_TC_DEBUG=2 HIDVAE_B200_LIB=$PWD/$lib timeout 300 python tools/bench_encode.py --tag $name-nogather --rows 4194304 --shape 32,256,3 --encode-only --reps 20 >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err;; esac
done
cat gpurun_out/ab.jsonl; tail -5 gpurun_out/ab.err
