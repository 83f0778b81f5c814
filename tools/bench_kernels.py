"""Per-kernel timings at the bench shapes (CUDA events, inputs larger than L2 or an L2 flush between launches): tuning runs
and A/B builds (HIDVAE_B200_LIB=hid-vae_b200/build/variants/X.so).

    python tools/bench_kernels.py [encoder] [rq] [train] [c4] [--reps 10] [--tag NAME]
"""
import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import torch  # noqa: E402

import bench  # noqa: E402
from hidvae_b200 import ops  # noqa: E402
from oracle import encoder as OE  # noqa: E402  (seeded weights only)

ap = argparse.ArgumentParser()
ap.add_argument("what", nargs="*", default=["encoder", "rq", "train", "c4"])
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--tag", default=os.path.basename(os.environ.get("HIDVAE_B200_LIB", "shipped")))
args = ap.parse_args()
torch.cuda.set_device(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
res = dict(tag=args.tag)


def ms(fn, fl=flush):
    return statistics.mean(bench.time_steps(fn, args.reps, 3, fl))


if "encoder" in args.what:
    n = 1 << 20
    image = ops.encoder_pack([w.cuda() for w in OE.seeded_weights(bench.DIMS, 2024)])
    x = bench.synth_items(n, 1, "cuda")
    z = torch.empty(n, 32, device="cuda")
    t = ms(lambda: ops.encoder_forward(x, image, normalize=True, out=z), None)     # 3.2 GB of input: nothing survives in L2
    res["encoder_1Mi"] = dict(ms=t, tflops=bench.FLOP_ENCODER * n / t / 1e9, gitems=n / t / 1e6)
    x16 = x.half()
    t16 = ms(lambda: ops.encoder_forward(x16, image, normalize=True, out=z), None)
    res["encoder_1Mi_fp16_items"] = dict(ms=t16, tflops=bench.FLOP_ENCODER * n / t16 / 1e9, gitems=n / t16 / 1e6)
    del x16
    for small in (128, 4096, 12101):
        xs = x[:small]
        res[f"encoder_{small}"] = dict(ms=ms(lambda: ops.encoder_forward(xs, image, normalize=True)))
    del x, z
if "rq" in args.what or "train" in args.what:
    n, d, k, L = 1 << 22, 32, 256, 3
    x, cbs, g_emb, g_loss = bench.synth_rq(n, d, k, L, 7, "cuda")
    packed = ops.pack_codebooks(cbs)
    if "rq" in args.what:
        t = ms(lambda: ops.rq_encode(x, cbs, packed=packed))
        res["rq_encode_4Mi"] = dict(ms=t, tflops=2.0 * k * d * L * n / t / 1e9, gitems=n / t / 1e6)
    if "train" in args.what:
        out = ops.rq_forward(x, cbs, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed)
        tf = ms(lambda: ops.rq_forward(x, cbs, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed))
        tb = ms(lambda: ops.rq_backward(x, cbs, out.ids, 3, True, 0.4, g_emb, g_loss, None))
        bf, bb = n * (4 * d + 4 * d * L + 8 * L + 4), n * (4 * d + 8 * L + 4 * d * L + 4 + 4 * d)
        ids_e = ops.rq_encode(x, cbs, packed=packed)
        res["bwd_ste_4Mi_ms"] = ms(lambda: ops.rq_backward(x, cbs, ids_e, ops.HV_MODE_STE, True, 0.4, g_emb, g_loss, None))
        res["train_4Mi"] = dict(fwd_ms=tf, bwd_ms=tb, fwd_gbs=bf / tf / 1e6, bwd_gbs=bb / tb / 1e6, both_gbs=(bf + bb) / (tf + tb) / 1e6)
    del x, g_emb
if "small" in args.what:
    for n in (1024, 12101, 32768):
        x, cbs, g_emb, g_loss = bench.synth_rq(n, 32, 256, 3, 7, "cuda")
        packed = ops.pack_codebooks(cbs)
        out = ops.rq_forward(x, cbs, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed)
        res[f"small_{n}"] = dict(
            enc_us=1e3 * ms(lambda: ops.rq_encode(x, cbs, packed=packed)),
            fwd_us=1e3 * ms(lambda: ops.rq_forward(x, cbs, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed)),
            bwd_us=1e3 * ms(lambda: ops.rq_backward(x, cbs, out.ids, 3, True, 0.4, g_emb, g_loss, None)))
if "c4" in args.what:
    n, d, k, L = 65536, 64, 4096, 4
    x, cbs, _g, _l = bench.synth_rq(n, d, k, L, 9, "cuda")
    packed = ops.pack_codebooks(cbs)
    t = ms(lambda: ops.rq_encode(x, cbs, packed=packed))
    res["c4"] = dict(ms=t, tflops=2.0 * k * d * L * n / t / 1e9, gitems=n / t / 1e6)
print(json.dumps(res))
