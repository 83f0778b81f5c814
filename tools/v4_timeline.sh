#!/bin/bash
# tools/v4_timeline.sh: clock64 timeline of block 0 of the streamed-image kernel at C4 (first epilogue warp and the MMA issuer), from
# the instrumented A/B build (tools/build_variant.sh v4_ts "-DHV_TC_INSTRUMENT" rq_fwd_tc_v4) -> gpurun_out/v4_ts.log
cd "$(dirname "$0")/.."
export HIDVAE_B200_LIB=$PWD/hid-vae_b200/build/variants/v4_ts.so
python - <<'PY' > gpurun_out/v4_ts.log 2>&1
import os, sys
sys.path[:0] = [os.path.join(os.getcwd(), "hid-vae_b200"), os.getcwd()]
import torch, bench
from hidvae_b200 import ops
x, cbs, _g, _l = bench.synth_rq(65536, 64, 4096, 4, 9, "cuda")
packed = ops.pack_codebooks(cbs)
for _ in range(2):
    ops.rq_encode(x, cbs, packed=packed)
torch.cuda.synchronize()
PY
grep -c "^TS" gpurun_out/v4_ts.log
python - <<'PY'
rows = [l.split() for l in open("gpurun_out/v4_ts.log") if l.startswith("TS")]
half = len(rows) // 2
rows = rows[half:]          # second launch
for who in (0, 1):
    e = [(int(c), int(t)) for _, w, c, t in rows if int(w) == who]
    if not e: continue
    t0 = e[0][1]
    print("who", who, " ".join(f"{c}@{(t - t0) & 0xffffffff}" for c, t in e[:140]))
PY
