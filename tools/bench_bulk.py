"""Bulk semantic-ID assignment (BASELINE.json config 5 shape, one GPU's share): `HSemanticIdTokenizer.precompute_corpus_ids`
on device-resident synthetic 768-d items -- encoder MLP (PyTorch / cuBLAS) + fused L-level quantiser -- next to the
quantiser alone on the 32-d encoder outputs.  SURVEY.md section 8d: 1,171,456 flop and 3,096 B per item end to end.

    python tools/bench_bulk.py [--items 2097152] [--reps 5] > gpurun_out/bulk.jsonl
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from hidvae_b200 import ops  # noqa: E402
from modules.tokenizer.h_semids import HSemanticIdTokenizer  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--items", type=int, default=1 << 21)
ap.add_argument("--reps", type=int, default=5)  # (the tokenizer option `encoder_tf32` makes the same choice per call)
args = ap.parse_args()
torch.cuda.set_device(0)
torch.manual_seed(0)
n = args.items
tok = HSemanticIdTokenizer(input_dim=768, output_dim=32, hidden_dims=[512, 256, 128], codebook_size=256, n_layers=3,
                           n_cat_feats=0, hrqvae_codebook_normalize=True, chunk_items=1 << 18).cuda().eval()
x = F.normalize(torch.randn(n, 768, device="cuda", generator=torch.Generator("cuda").manual_seed(1)), dim=-1)


def ev_time(fn, reps, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


with torch.no_grad():
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        t_all = ev_time(lambda: tok.precompute_corpus_ids(x), args.reps)
        t_enc = ev_time(lambda: [tok.hrq_vae.encode(x[i:i + (1 << 18)]) for i in range(0, n, 1 << 18)], args.reps)
        print(json.dumps(dict(case="bulk_assign_768d", items=n, encoder_matmul="tf32" if tf32 else "fp32", ms=t_all,
                              items_per_s=n / t_all * 1e3, tflops=1171456.0 * n / t_all / 1e9, hbm_gbs=3096.0 * n / t_all / 1e6,
                              encoder_only_ms=t_enc)))
    enc = tok.hrq_vae.encode(x[: 1 << 20])
    cbs = tok.hrq_vae.effective_codebooks().detach()
    packed = ops.pack_codebooks(cbs)
    big = enc.repeat(4, 1)
    t_rq = ev_time(lambda: ops.rq_encode(big, cbs, packed=packed), args.reps)
    print(json.dumps(dict(case="rq_only_32d", items=big.shape[0], ms=t_rq, items_per_s=big.shape[0] / t_rq * 1e3,
                          tflops=49152.0 * big.shape[0] / t_rq / 1e9)))
