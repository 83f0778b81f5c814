#!/bin/bash
# clock64 timeline of warpgroup 0 of block 0, TRAINING forward (rotation trick, emb_out + loss), instrumented A/B build v11_ts
cd "$(dirname "$0")/.."
export HIDVAE_B200_LIB=$PWD/hid-vae_b200/build/variants/v11_ts.so
export HIDVAE_TC_DEBUG=64
python - <<'PY' > gpurun_out/v11_ts_train.log 2>&1
import os, sys
sys.path[:0] = [os.path.join(os.getcwd(), "hid-vae_b200"), os.getcwd()]
import torch, bench
from hidvae_b200 import ops
x, cbs, _g, _l = bench.synth_rq(1 << 20, 32, 256, 3, 7, "cuda")
packed = ops.pack_codebooks(cbs)
for _ in range(2):
    ops.rq_forward(x, cbs, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed)
torch.cuda.synchronize()
PY
python tools/ts_show.py gpurun_out/v11_ts_train.log | grep -v "^who" | tail -24
