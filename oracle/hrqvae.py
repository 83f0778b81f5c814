"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the untagged `HRqVae.forward` step and of the reference's bulk
assignment loop, built from oracle/encoder.py and oracle/rq.py.  Used as the CPU baseline legs of bench.py
 and by tests; never by the product path.

  untagged_step           /root/reference/modules/h_rqvae.py:585-640 with batch.tags_* = None: encode -> L quantiser
                          levels -> decode(sum of level embeddings) -> SSE reconstruction -> loss.mean() + rq.mean()
                          (tag losses and the uniqueness term are 0 without tags / as wired)
  precompute_corpus_ids   /root/reference/modules/tokenizer/h_semids.py:109-131: DataLoader batches of 512 ->
                          encode -> get_semantic_ids (eval) -> concatenated ids
"""
from typing import Sequence

import torch
from torch import Tensor

from oracle import encoder as OE
from oracle import rq as O


def untagged_step(x: Tensor, enc_w: Sequence[Tensor], dec_w: Sequence[Tensor], codebook_w: Sequence[Tensor], mode: int,
                  beta: float, normalize: bool, n_cat_feats: int = 0) -> Tensor:
    """Scalar training loss of one untagged step (autograd-tracked through every argument that requires grad)."""
    enc = OE.mlp_forward(x, enc_w, normalize)                                   # h_rqvae.py:599 (encoder.py:23-36)
    cbs = [O.effective_codebook(w, normalize and l == 0) for l, w in enumerate(codebook_w)]   # h_rqvae.py:295
    q = O.rq_forward(enc, cbs, mode, beta, True)                                # h_rqvae.py:602 (:515-574)
    x_hat = OE.mlp_forward(q.embeddings.sum(dim=-1), dec_w, False)              # h_rqvae.py:607
    if n_cat_feats:                                                             # h_rqvae.py:610 (identity when n_cat == 0:
        x_hat = torch.cat([torch.nn.functional.normalize(x_hat[..., :-n_cat_feats], dim=-1),   # x_hat[..., :-0] is empty)
                           x_hat[..., -n_cat_feats:]], dim=-1)
    recon = ((x_hat - x) ** 2).sum(dim=-1)                                      # loss.py:11-12
    return recon.mean() + q.quantize_loss.mean()                                # h_rqvae.py:633-639


@torch.no_grad()
def precompute_corpus_ids(x: Tensor, enc_w: Sequence[Tensor], codebooks: Sequence[Tensor], normalize: bool,
                          batch_size: int = 512) -> Tensor:
    """[N, L] int64 ids of a catalogue, walked in the reference's batches of 512 (h_semids.py:119)."""
    out = []
    for lo in range(0, x.shape[0], batch_size):
        enc = OE.mlp_forward(x[lo:lo + batch_size], enc_w, normalize)
        out.append(O.rq_forward(enc, codebooks, O.MODE_STE, 0.25, False).sem_ids)
    return torch.cat(out)
