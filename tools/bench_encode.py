"""Encode / train-forward / backward timings for one shape (CUDA events, L2 flushed), used for tuning runs.

    [HIDVAE_TC_NWG=2|4] python tools/bench_encode.py --rows 4194304 --shape 32,256,3 [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import torch  # noqa: E402

import bench  # noqa: E402
from hidvae_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1 << 22)
ap.add_argument("--shape", default="32,256,3")
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--tag", default="")
ap.add_argument("--encode-only", action="store_true")
ap.add_argument("--full", action="store_true")
args = ap.parse_args()
d, k, L = (int(v) for v in args.shape.split(","))
torch.cuda.set_device(0)
x, cbs, g_emb, g_loss = bench.synth(args.rows, d, k, L, 7, "cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
packed = ops.pack_codebooks(cbs)
res = dict(tag=args.tag, impl=os.environ.get("HIDVAE_TC_IMPL", "v5"), rows=args.rows, d=d, k=k, L=L)
t = bench.time_region(lambda: ops.rq_encode(x, cbs, packed=packed), args.reps, 3, flush) / args.reps
res.update(encode_ms=t, encode_gitems=args.rows / t / 1e6, encode_tflops=2.0 * k * d * L * args.rows / t / 1e9)
if d <= 32 and not args.encode_only:
    out = ops.rq_forward(x, cbs, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed)
    tf = bench.time_region(lambda: ops.rq_forward(x, cbs, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed), args.reps, 3, flush) / args.reps
    tb = bench.time_region(lambda: ops.rq_backward(x, cbs, out.ids, 3, True, 0.4, g_emb, g_loss, None), args.reps, 3, flush) / args.reps
    byt = args.rows * (4 * d * (3 + 2 * L) + 16 * L + 8)
    res.update(train_fwd_ms=tf, train_bwd_ms=tb, bwd_gbs=args.rows * (4 * d * (2 + L) + 8 * L + 4) / tb / 1e6, train_gbs=byt / (tf + tb) / 1e6)
print(json.dumps(res))
