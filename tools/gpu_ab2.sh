#!/bin/bash
# small-N A/B with more repetitions (variant libraries under hid-vae_b200/build/variants)
mkdir -p gpurun_out; rm -f gpurun_out/ab2.jsonl
for rep in 1 2; do for lib in hid-vae_b200/build/variants/*.so; do
  name=$(basename $lib .so)
  HIDVAE_B200_LIB=$PWD/$lib timeout 300 python tools/bench_encode.py --tag $name --rows 12101 --shape 32,256,3 --reps 200 >> gpurun_out/ab2.jsonl 2>/dev/null
done; done
python - <<'PY'
import json
for l in open('gpurun_out/ab2.jsonl'):
    d=json.loads(l); print(d['tag'],d['rows'],'enc %.4f'%d['encode_ms'],'trainfwd %.4f'%d.get('train_fwd_ms',0),'bwd %.4f'%d.get('train_bwd_ms',0))
PY
