#!/bin/bash
# tools/ab_rq.sh [variant ...]: parity tests of the quantiser on the shipped build, then the 4 Mi-row encode timing of the shipped
# build and of every named A/B build (hid-vae_b200/build/variants/NAME.so), into gpurun_out/ab_rq.jsonl
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_rq.py tests/test_gpu_modules.py -x -q -m gpu > gpurun_out/ab_pytest.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/ab_pytest.log
: > gpurun_out/ab_rq.jsonl
timeout 120 python tools/bench_kernels.py rq train --tag shipped >> gpurun_out/ab_rq.jsonl 2>gpurun_out/ab_rq.err
for v in "$@"; do
  HIDVAE_B200_LIB=$PWD/hid-vae_b200/build/variants/$v.so timeout 120 python tools/bench_kernels.py rq train --tag $v >> gpurun_out/ab_rq.jsonl 2>>gpurun_out/ab_rq.err
done
cat gpurun_out/ab_rq.jsonl; tail -5 gpurun_out/ab_rq.err
