"""Short driver for ncu: a few C2 steps (train fwd + bwd + encode) and a few large encode-only launches.

    python tools/profile_step.py [--rows 1048576] [--reps 3]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import torch  # noqa: E402

import bench  # noqa: E402
from hidvae_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--shape", default="32,256,3")
args = ap.parse_args()
d, k, L = (int(v) for v in args.shape.split(","))

torch.cuda.set_device(0)
w = bench.WORKLOAD
x, cbs, g_emb, g_loss = bench.synth(w["n_items"], w["embed_dim"], w["codebook_size"], w["n_levels"], 0, "cuda")
step = bench.NativeStep(ops, x, cbs, g_emb, g_loss, w["beta"])
for _ in range(args.reps):
    step()
xb, cb, ge, gl = bench.synth(args.rows, d, k, L, 1, "cuda")
packed = ops.pack_codebooks(cb)
for _ in range(args.reps):
    ids = ops.rq_encode(xb, cb, packed=packed)
if d == 32:
    for _ in range(args.reps):
        out = ops.rq_forward(xb, cb, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed)
        ops.rq_backward(xb, cb, out.ids, 3, True, 0.4, ge, gl, None)
torch.cuda.synchronize()
print("ok", int(ids.sum()))
