#!/bin/bash
# tools/build_variant.sh NAME "EXTRA nvcc flags" [source.cu]: an A/B build of one kernel file, linked with the other
# objects of the current build into hid-vae_b200/build/variants/NAME.so (select it with HIDVAE_B200_LIB=...; tools/bench_kernels.py times it).
set -e
cd "$(dirname "$0")/../hid-vae_b200"
name=$1; extra=$2; src=${3:-rq_fwd_tc_v11}
make -s
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $extra -c csrc/$src.cu -o build/variants/$name.o
objs=$(ls build/*.o | grep -v "/$src.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $objs build/variants/$name.o
rm build/variants/$name.o
echo built build/variants/$name.so
