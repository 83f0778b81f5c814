#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ablate.jsonl
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for dbg in 0 4 1; do
  HIDVAE_BWD_DEBUG=$dbg timeout 300 python tools/bench_encode.py --tag bwd$dbg --rows 4194304 --shape 32,256,3 >> gpurun_out/ablate.jsonl 2>> gpurun_out/ablate.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ablate.jsonl'):
    d=json.loads(l); print(d['tag'], 'enc', round(d['encode_ms'],3), 'fwd', round(d['train_fwd_ms'],3), 'bwd', round(d['train_bwd_ms'],3))
PY
tail -3 gpurun_out/ablate.err
