"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- CPU restatement of HiD-VAE's residual quantiser.

Every function cites the reference lines it follows (paths relative to the reference root).  The same ATen
CPU operators the reference uses (`mm`, `min`, `normalize`, `bmm`-shaped products) are used on purpose, in
the same order, so that on a CPU the oracle reproduces the reference to the last bit for the forward pass
and autograd reproduces its gradients.  Pinned by tests/golden (made from the reference by make_golden.py).
"""
from __future__ import annotations

from typing import List, NamedTuple, Optional, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor

# numeric values follow QuantizeForwardMode (modules/quantize.py:17-20)
MODE_GUMBEL_SOFTMAX = 1
MODE_STE = 2
MODE_ROTATION_TRICK = 3


class LevelOutput(NamedTuple):
    embeddings: Tensor  # emb_out [N, D]
    ids: Tensor         # [N] int64
    loss: Tensor        # [N]


class RqOutput(NamedTuple):
    embeddings: Tensor     # [N, D, L]
    residuals: Tensor      # [N, D, L]  (the input of every level)
    sem_ids: Tensor        # [N, L] int64
    quantize_loss: Tensor  # [N]
    level_losses: List[Tensor]


def effective_codebook(weight: Tensor, normalize: bool, sim_vq_weight: Optional[Tensor] = None) -> Tensor:
    """`out_proj(embedding.weight)`: optional bias-free Linear then optional row L2 norm.
    modules/quantize.py:70-73,106; modules/normalize.py:7-8 (eps 1e-12)."""
    cb = weight
    if sim_vq_weight is not None:
        cb = F.linear(cb, sim_vq_weight)
    if normalize:
        cb = F.normalize(cb, p=2, dim=-1, eps=1e-12)
    return cb


def squared_l2_table(x: Tensor, codebook: Tensor) -> Tensor:
    """[N, K] table  |x|^2 + |c|^2 - 2 x.c  with the reference's operator order.  modules/quantize.py:108-113."""
    xx = (x ** 2).sum(axis=1, keepdim=True)
    cc = (codebook.T ** 2).sum(axis=0, keepdim=True)
    return xx + cc - 2 * x @ codebook.T


def nearest_code(x: Tensor, codebook: Tensor) -> Tensor:
    """argmin over the table; `min` returns the first index among exact ties.  modules/quantize.py:122."""
    return squared_l2_table(x, codebook).detach().min(axis=1).indices


def quantize_loss(query: Tensor, value: Tensor, beta: float) -> Tensor:
    """codebook term + beta * commitment term, per row.  modules/loss.py:41-44."""
    codebook_term = ((query.detach() - value) ** 2).sum(axis=[-1])
    commit_term = ((query - value.detach()) ** 2).sum(axis=[-1])
    return codebook_term + beta * commit_term


def rotation_trick(u: Tensor, q: Tensor, e: Tensor) -> Tensor:
    """e - 2 (e.w) w + 2 (e.u) q with u, q, w detached.  modules/quantize.py:34-45 (incl. the final squeeze)."""
    w = F.normalize(u + q, p=2, dim=1, eps=1e-6).detach()
    e3 = e.unsqueeze(1)                                          # [N, 1, D]
    refl = e3 @ w.unsqueeze(2) @ w.unsqueeze(1)                 # (e.w) w
    rot = e3 @ u.unsqueeze(2).detach() @ q.unsqueeze(1).detach()  # (e.u) q
    return (e3 - 2 * refl + 2 * rot).squeeze()


def gumbel_weights(neg_dist: Tensor, temperature: float, uniform: Tensor, eps: float = 1e-20) -> Tensor:
    """softmax((logits + G) / T), G = -log(-log(U + eps) + eps).  distributions/gumbel.py:8-18."""
    g = -torch.log(-torch.log(uniform + eps) + eps)
    return F.softmax((neg_dist + g) / temperature, dim=-1)


def quantize_level(
    x: Tensor,
    codebook: Tensor,
    mode: int,
    beta: float,
    training: bool,
    temperature: float = 0.2,
    uniform: Optional[Tensor] = None,
) -> LevelOutput:
    """One `Quantize.forward` given the *effective* codebook.  modules/quantize.py:100-154."""
    assert x.shape[-1] == codebook.shape[-1]
    dist = squared_l2_table(x, codebook)
    ids = dist.detach().min(axis=1).indices
    if training:
        if mode == MODE_GUMBEL_SOFTMAX:
            if uniform is None:
                uniform = torch.rand(dist.shape)
            emb = gumbel_weights(-dist, temperature, uniform) @ codebook
            emb_out = emb
        elif mode == MODE_STE:
            emb = codebook[ids]
            emb_out = x + (emb - x).detach()
        elif mode == MODE_ROTATION_TRICK:
            emb = codebook[ids]
            emb_out = rotation_trick(
                x / (x.norm(dim=-1, keepdim=True) + 1e-8),
                emb / (emb.norm(dim=-1, keepdim=True) + 1e-8),
                x,
            )
        else:
            raise Exception("Unsupported Quantize forward mode.")
        loss = quantize_loss(x, emb, beta)
    else:
        emb_out = codebook[ids]
        loss = quantize_loss(x, emb_out, beta)
    return LevelOutput(emb_out, ids, loss)


def rq_forward(
    enc: Tensor,
    codebooks: Sequence[Tensor],
    mode: int,
    beta: float,
    training: bool,
    temperature: float = 0.2,
) -> RqOutput:
    """The residual loop of `HRqVae.get_semantic_ids` without the tag heads (they only read emb_out).
    modules/h_rqvae.py:500,515-523,552,572-574."""
    res = enc
    total = torch.tensor(0.0)
    embs, residuals, ids, level_losses = [], [], [], []
    for cb in codebooks:
        residuals.append(res)
        out = quantize_level(res, cb, mode, beta, training, temperature)
        total = total + out.loss
        level_losses.append(out.loss)
        embs.append(out.embeddings)
        ids.append(out.ids)
        res = res - out.embeddings
    return RqOutput(
        embeddings=torch.stack(embs, 0).permute(1, 2, 0),
        residuals=torch.stack(residuals, 0).permute(1, 2, 0),
        sem_ids=torch.stack(ids, 0).permute(1, 0),
        quantize_loss=total,
        level_losses=level_losses,
    )


def rq_backward_closed_form(
    enc: Tensor,
    codebooks: Sequence[Tensor],
    sem_ids: Tensor,
    g_emb: Tensor,
    g_loss: Tensor,
    mode: int,
    beta: float,
):
    """The backward recursion the CUDA kernel implements ("backward"), in float64-free
    plain tensor algebra, *without* autograd.  It is itself checked against autograd through `rq_forward`
    (tests/test_oracle_golden.py) so the kernel can be compared with either.

    g_emb [N, D, L] is dLoss/d embeddings, g_loss [N] is dLoss/d quantize_loss.
    Returns (g_enc [N, D], g_codebooks list of [K, D])."""
    L = len(codebooks)
    res = [enc]
    es, us, qs, ws = [], [], [], []
    r = enc
    for l in range(L):
        e = codebooks[l][sem_ids[:, l]]
        es.append(e)
        if mode == MODE_ROTATION_TRICK:
            u = r / (r.norm(dim=-1, keepdim=True) + 1e-8)
            q = e / (e.norm(dim=-1, keepdim=True) + 1e-8)
            w = F.normalize(u + q, p=2, dim=1, eps=1e-6)
            o = r - 2 * (r * w).sum(-1, keepdim=True) * w + 2 * (r * u).sum(-1, keepdim=True) * q
            us.append(u), qs.append(q), ws.append(w)
        elif mode == MODE_STE:
            o = e
        else:
            raise Exception("closed form exists for STE and ROTATION_TRICK only")
        r = r - o
        res.append(r)
    G = torch.zeros_like(enc)
    g_cbs = [torch.zeros_like(cb) for cb in codebooks]
    gl = g_loss.unsqueeze(-1)
    for l in reversed(range(L)):
        h = g_emb[:, :, l] - G
        if mode == MODE_ROTATION_TRICK:
            u, q, w = us[l], qs[l], ws[l]
            jh = h - 2 * (h * w).sum(-1, keepdim=True) * w + 2 * (h * q).sum(-1, keepdim=True) * u
        else:
            jh = h
        G = G + jh + 2 * beta * (res[l] - es[l]) * gl
        g_cbs[l].index_add_(0, sem_ids[:, l], 2 * (es[l] - res[l]) * gl)
    return G, g_cbs


def uniqueness_loss(sem_ids: Tensor, feats: Tensor, margin: float, weight: float) -> Tensor:
    """weight * mean over pairs i<j with identical id rows of relu(cos(f_i, f_j) - margin); 0 when there is
    no such pair or fewer than two rows.  modules/h_rqvae.py:41-105.  `sem_ids` is [rows, width]; the
    reference's own call site passes the transposed [L, B] tensor (h_rqvae.py:630-631) -- callers choose."""
    n_rows = sem_ids.shape[0]
    zero = torch.tensor(0.0)
    if n_rows <= 1:
        return zero
    same = (sem_ids.unsqueeze(1) == sem_ids.unsqueeze(0)).all(dim=-1)
    same = same & ~torch.eye(n_rows, dtype=torch.bool)
    if not same.any():
        return zero
    a, b = torch.where(same)
    keep = a < b
    a, b = a[keep], b[keep]
    if len(a) == 0:
        return zero
    fa = F.normalize(feats[a], p=2, dim=-1)
    fb = F.normalize(feats[b], p=2, dim=-1)
    hinge = F.relu((fa * fb).sum(dim=-1) - margin)
    return weight * hinge.mean()


def p_unique_ids(sem_ids: Tensor) -> Tensor:
    """fraction of rows with no *later* identical row.  modules/h_rqvae.py:645-648 (sem_ids [N, L])."""
    same = (sem_ids.unsqueeze(1) == sem_ids.unsqueeze(0)).all(axis=-1)
    return (~torch.triu(same, diagonal=1)).all(axis=1).sum() / sem_ids.shape[0]


def embs_norm(embeddings: Tensor) -> Tensor:
    """`embs.norm(dim=1)` on [N, D, L] -> [N, L].  modules/h_rqvae.py:645."""
    return embeddings.norm(dim=1)
