#!/bin/bash
# parity of the tcgen05 forward (default build), then encode / train timings of generation 10 against the older ones
mkdir -p gpurun_out; rm -f gpurun_out/v10*.log gpurun_out/v10.jsonl gpurun_out/v10.err
timeout 600 python -m pytest tests/test_gpu_rq.py tests/test_gpu_modules.py -m gpu -q -x > gpurun_out/v10_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/v10_pytest.log
for impl in v10 v7; do
  for rows in 4194304 1048576 262144 65536 12101; do
  HIDVAE_TC_IMPL=$impl timeout 300 python tools/bench_encode.py --tag $impl --rows $rows --shape 32,256,3 --reps 20 >> gpurun_out/v10.jsonl 2>> gpurun_out/v10.err
  done
done
cat gpurun_out/v10.jsonl; tail -5 gpurun_out/v10.err
