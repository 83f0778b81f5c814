#!/bin/bash
# backward parity + timing (4 Mi rows and the C2 shape)
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x -k "backward or module or golden" 2>&1 | tail -3
for rows in 4194304 12101; do timeout 300 python tools/bench_encode.py --tag bwd --rows $rows --shape 32,256,3 --reps 20 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(d['rows'],'bwd %.4f ms'%d['train_bwd_ms'],'fwd %.4f'%d['train_fwd_ms'],'enc %.4f'%d['encode_ms'])"; done
