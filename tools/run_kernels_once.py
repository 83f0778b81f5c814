"""A handful of launches of every hot kernel at its bench shape, for `ncu` captures (tools/ncu_capture.sh):
    python tools/run_kernels_once.py [encoder|rq|train|c4|aux|all]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "hid-vae_b200"), ROOT]
import torch
import torch.nn.functional as F
from hidvae_b200 import ops
from oracle import encoder as OE

what = sys.argv[1] if len(sys.argv) > 1 else "all"
torch.manual_seed(0)
if what in ("encoder", "all"):
    n = 1 << 20
    ws = [w.cuda() for w in OE.seeded_weights([768, 512, 256, 128, 32], 2024)]
    image = ops.encoder_pack(ws)
    x = F.normalize(torch.randn(n, 768, device="cuda"), dim=-1)
    z = torch.empty(n, 32, device="cuda")
    for _ in range(3):
        ops.encoder_forward(x, image, normalize=True, out=z)
    torch.cuda.synchronize()
    del x
if what in ("rq", "train", "all"):
    n, d, k, L = 1 << 22, 32, 256, 3
    x = F.normalize(torch.randn(n, d, device="cuda"), dim=-1)
    cbs = torch.rand(L, k, d, device="cuda")
    cbs[0] = F.normalize(cbs[0], dim=-1)
    for l in range(1, L):
        cbs[l] = (cbs[l] - 0.5) * (0.7 * 0.5 ** l)
    packed = ops.pack_codebooks(cbs)
    if what in ("rq", "all"):
        for _ in range(3):
            ops.rq_encode(x, cbs, packed=packed)
    if what in ("train", "all"):
        g_emb = torch.randn(L, n, d, device="cuda") * 0.01
        g_loss = torch.full((n,), 1.0 / n, device="cuda")
        for _ in range(2):
            out = ops.rq_forward(x, cbs, 3, True, 0.4, want_emb=True, want_loss=True, packed=packed)
            ops.rq_backward(x, cbs, out.ids, 3, True, 0.4, g_emb, g_loss, None)
    torch.cuda.synchronize()
    del x
if what in ("c4", "all"):
    n, d, k, L = 65536, 64, 4096, 4
    x = F.normalize(torch.randn(n, d, device="cuda"), dim=-1)
    cbs = torch.rand(L, k, d, device="cuda")
    packed = ops.pack_codebooks(cbs)
    for _ in range(3):
        ops.rq_encode(x, cbs, packed=packed)
    torch.cuda.synchronize()
if what in ("aux", "all"):
    # stress shapes of the k-means update (segmented form) and the uniqueness loss (sorted form)
    n, d, k = 65536, 64, 4096
    x = F.normalize(torch.randn(n, d, device="cuda"), dim=-1)
    cent = x[torch.randperm(n, device="cuda")[:k]].clone()
    for _ in range(2):
        assign = ops.kmeans_assign(x, cent, exact_diff_form=False)
        sums, counts, _ = ops.kmeans_accumulate(x, assign, k)
        ops.kmeans_finalize(sums, counts, cent.clone())
    ids = torch.randint(0, 24, (n, 3), device="cuda")
    feats = torch.randn(n, 32, device="cuda", requires_grad=True)
    for _ in range(2):
        ops.uniqueness_loss(ids, feats, 0.05, 1.0).backward()
    torch.cuda.synchronize()
print("done", what)
