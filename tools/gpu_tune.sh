#!/bin/bash
# quick A/B of kernel variants (no ncu): parity subset + timings
mkdir -p gpurun_out; rm -f gpurun_out/tune.jsonl gpurun_out/status.txt
timeout 900 python -m pytest tests/test_gpu_rq.py -m gpu -q -x > gpurun_out/pytest_rq.log 2>&1; echo "pytest_rq rc=$?" >> gpurun_out/status.txt
for impl in v5 ${TUNE_V4:+v4}; do
  HIDVAE_TC_IMPL=$impl timeout 300 python tools/bench_encode.py --tag $impl --rows 4194304 --shape 32,256,3 >> gpurun_out/tune.jsonl 2>> gpurun_out/tune.err
  HIDVAE_TC_IMPL=$impl timeout 300 python tools/bench_encode.py --tag $impl --rows 12101 --shape 32,256,3 --reps 50 >> gpurun_out/tune.jsonl 2>> gpurun_out/tune.err
  HIDVAE_TC_IMPL=$impl timeout 300 python tools/bench_encode.py --tag $impl --rows 65536 --shape 64,4096,4 >> gpurun_out/tune.jsonl 2>> gpurun_out/tune.err
done
cat gpurun_out/status.txt gpurun_out/tune.jsonl; tail -15 gpurun_out/pytest_rq.log; tail -5 gpurun_out/tune.err
