#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/ablate.jsonl
timeout 900 python -m pytest tests -m gpu -q -x -k "backward" 2>&1 | tail -3
for dbg in 0; do
  HIDVAE_BWD_DEBUG=$dbg timeout 300 python tools/bench_encode.py --tag bwd$dbg --rows 4194304 --shape 32,256,3 >> gpurun_out/ablate.jsonl 2>> gpurun_out/ablate.err
  HIDVAE_BWD_DEBUG=$dbg timeout 300 python tools/bench_encode.py --tag bwd$dbg --rows 12101 --shape 32,256,3 --reps 50 >> gpurun_out/ablate.jsonl 2>> gpurun_out/ablate.err
done
python - <<'PY'
import json
for l in open('gpurun_out/ablate.jsonl'):
    d=json.loads(l); print(d['tag'], d['rows'], 'enc', round(d['encode_ms'],4), 'fwd', round(d['train_fwd_ms'],4), 'bwd', round(d['train_bwd_ms'],4))
PY
tail -3 gpurun_out/ablate.err
