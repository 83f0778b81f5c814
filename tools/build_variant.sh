#!/bin/bash
# tools/build_variant.sh NAME "EXTRA nvcc flags" ["source1 source2 ..."]: an A/B build of some kernel files, linked with the other
# objects of the current build into hid-vae_b200/build/variants/NAME.so (select it with HIDVAE_B200_LIB=...; tools/bench_kernels.py times it).
set -e
cd "$(dirname "$0")/../hid-vae_b200"
name=$1; extra=$2; srcs=${3:-rq_fwd_tc_v11}
make -s
mkdir -p build/variants
objs=$(ls build/*.o)
vobjs=""
for src in $srcs; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC $extra -c csrc/$src.cu -o build/variants/$name.$src.o
  objs=$(echo "$objs" | grep -v "/$src.o")
  vobjs="$vobjs build/variants/$name.$src.o"
done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build/variants/$name.so $objs $vobjs
rm $vobjs
echo built build/variants/$name.so
