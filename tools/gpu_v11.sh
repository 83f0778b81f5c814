#!/bin/bash
# parity of the tcgen05 forward with generation 11 forced, then encode / train timings against the older generations
mkdir -p gpurun_out; rm -f gpurun_out/v11*.log gpurun_out/v11.jsonl gpurun_out/v11.err
HIDVAE_TC_IMPL=v11 timeout 600 python -m pytest tests/test_gpu_rq.py -m gpu -q -x > gpurun_out/v11_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/v11_pytest.log
for impl in v11 ${V11_AB:-v7}; do
  for rows in 4194304 262144 12101; do
  HIDVAE_TC_IMPL=$impl timeout 300 python tools/bench_encode.py --tag $impl --rows $rows --shape 32,256,3 --reps 20 >> gpurun_out/v11.jsonl 2>> gpurun_out/v11.err
  done
done
cat gpurun_out/v11.jsonl; tail -5 gpurun_out/v11.err
