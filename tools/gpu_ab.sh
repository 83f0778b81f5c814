#!/bin/bash
# A/B timing of the variant libraries under hid-vae_b200/build/variants (encode + train forward, three sizes)
mkdir -p gpurun_out; rm -f gpurun_out/ab.jsonl gpurun_out/ab.err
for lib in hid-vae_b200/build/variants/*.so; do
  name=$(basename $lib .so)
  if [ -n "$AB_TEST" ]; then HIDVAE_B200_LIB=$PWD/$lib timeout 600 python -m pytest tests/test_gpu_rq.py -m gpu -q -x 2>&1 | tail -2; fi
  for rows in 4194304 262144 12101; do
    HIDVAE_B200_LIB=$PWD/$lib timeout 300 python tools/bench_encode.py --tag $name --rows $rows --shape 32,256,3 --reps 20 >> gpurun_out/ab.jsonl 2>> gpurun_out/ab.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/ab.jsonl'):
    d=json.loads(l); print(d['tag'],d['rows'],'enc %.4f'%d['encode_ms'],'trainfwd %.4f'%d.get('train_fwd_ms',0))
PY
tail -5 gpurun_out/ab.err
