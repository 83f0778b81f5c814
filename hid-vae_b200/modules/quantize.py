"""One residual-quantisation level with the module API of the reference's modules/quantize.py
(`Quantize`, `QuantizeForwardMode`, `QuantizeDistance`, `QuantizeOutput`), computed by the sm_100a kernels of
libhidvae_b200.so.

What the reference does with ~20 ATen launches per level -- the [N, K] distance table (quantize.py:109-113),
argmin (:122), gather (:97-98), STE / rotation-trick value (:131-140), QuantizeLoss (:144, :148) -- is ONE
fused kernel here (forward) and one fused kernel in backward; the [N, K] table never exists in memory.  The
effective codebook `out_proj(embedding.weight)` (row L2 norm and/or the sim_vq Linear, :70-73, :106) stays in
PyTorch: it is K x D, and autograd carries the kernel's codebook gradient back through it.

State-dict keys (`embedding.weight`, `out_proj.0.weight`) equal the reference's, so its checkpoints load.
GUMBEL_SOFTMAX (the class default; training mode :125-130) is one fused kernel per level and direction as well
(`hv_gumbel_forward/backward`: distance, in-kernel Philox noise, softmax and the soft gather `w @ codebook` without any
[N, K] tensor in memory); its eval mode is the same fused kernel as STE.  Only the COSINE distance (selected by nothing
in HiD-VAE) and Gumbel shapes outside D in {16, 32, 64}, K <= 256 run on PyTorch GPU ops with the reference's formulas.
CPU tensors are rejected: there is no CPU path.
"""
from enum import Enum
from typing import NamedTuple, Optional

import torch
from torch import Tensor, nn

from distributions.gumbel import gumbel_softmax_sample
from hidvae_b200 import ops
from hidvae_b200.gin_lite import constants_from_enum
from init.kmeans import kmeans_init_
from modules.loss import QuantizeLoss
from modules.normalize import L2NormalizationLayer


@constants_from_enum(module="modules.quantize")
class QuantizeForwardMode(Enum):
    GUMBEL_SOFTMAX = 1
    STE = 2
    ROTATION_TRICK = 3


class QuantizeDistance(Enum):
    L2 = 1
    COSINE = 2


class QuantizeOutput(NamedTuple):
    embeddings: Tensor
    ids: Tensor
    loss: Tensor


def efficient_rotation_trick_transform(u: Tensor, q: Tensor, e: Tensor) -> Tensor:
    """e - 2 (e.w) w + 2 (e.u) q with w = normalize(u + q) (reference quantize.py:34-45), on PyTorch ops.
    The fused kernel evaluates the same expression per row in registers; this stand-alone function exists for API
    compatibility and as the GPU-side cross-check in the tests."""
    w = torch.nn.functional.normalize(u + q, p=2, dim=1, eps=1e-6).detach()
    u, q = u.detach(), q.detach()
    ew = (e * w).sum(dim=1, keepdim=True)
    eu = (e * u).sum(dim=1, keepdim=True)
    return (e - 2 * ew * w + 2 * eu * q).unsqueeze(1).squeeze()


class Quantize(nn.Module):
    def __init__(self, embed_dim: int, n_embed: int, do_kmeans_init: bool = True, codebook_normalize: bool = False,
                 sim_vq: bool = False, commitment_weight: float = 0.25,
                 forward_mode: QuantizeForwardMode = QuantizeForwardMode.GUMBEL_SOFTMAX,
                 distance_mode: QuantizeDistance = QuantizeDistance.L2) -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.n_embed = n_embed
        self.embedding = nn.Embedding(n_embed, embed_dim)
        self.forward_mode = forward_mode
        self.distance_mode = distance_mode
        self.do_kmeans_init = do_kmeans_init
        self.kmeans_initted = False
        self.kmeans_process_group = None  # set by a data-parallel trainer so that all ranks get the same centroids
        self.algo = "auto"                # hv_algo_t selector ("auto" | "tcgen05" | "simt")
        self.out_proj = nn.Sequential(
            nn.Linear(embed_dim, embed_dim, bias=False) if sim_vq else nn.Identity(),
            L2NormalizationLayer(dim=-1) if codebook_normalize else nn.Identity(),
        )
        self.quantize_loss = QuantizeLoss(commitment_weight)
        self._init_weights()

    # ---- reference surface ------------------------------------------------------------------------------------
    @property
    def weight(self) -> Tensor:
        return self.embedding.weight

    @property
    def device(self) -> torch.device:
        return self.embedding.weight.device

    @property
    def commitment_weight(self) -> float:
        return self.quantize_loss.commitment_weight

    def _init_weights(self) -> None:
        nn.init.uniform_(self.embedding.weight)  # reference :86-89

    @torch.no_grad()
    def _kmeans_init(self, x: Tensor) -> None:
        kmeans_init_(self.embedding.weight, x=x, process_group=self.kmeans_process_group)
        self.kmeans_initted = True

    def get_item_embeddings(self, item_ids: Tensor) -> Tensor:
        return self.out_proj(self.embedding(item_ids))

    def effective_codebook(self) -> Tensor:
        """`out_proj(embedding.weight)` [K, D] -- what the distance, the gather and the loss all see (reference :106)."""
        return self.out_proj(self.embedding.weight)

    @property
    def fused(self) -> bool:
        """True when this level runs on the fused kernels (STE / rotation trick with the L2 distance)."""
        return (self.distance_mode == QuantizeDistance.L2
                and self.forward_mode in (QuantizeForwardMode.STE, QuantizeForwardMode.ROTATION_TRICK))

    def needs_kmeans(self) -> bool:
        return self.do_kmeans_init and not self.kmeans_initted

    # ---- forward ------------------------------------------------------------------------------------------------
    def forward(self, x: Tensor, temperature: float) -> QuantizeOutput:
        assert x.shape[-1] == self.embed_dim
        if not x.is_cuda:
            raise RuntimeError("Quantize.forward: CUDA tensors only (hidvae_b200 has no CPU fallback)")
        if self.needs_kmeans():
            self._kmeans_init(x=x)
        codebook = self.effective_codebook()
        if self.distance_mode not in (QuantizeDistance.L2, QuantizeDistance.COSINE):
            raise Exception("Unsupported Quantize distance mode.")
        if self.forward_mode not in tuple(QuantizeForwardMode):
            raise Exception("Unsupported Quantize forward mode.")

        if self.fused or (not self.training and self.distance_mode == QuantizeDistance.L2):
            # eval semantics are mode-independent (emb_out = codebook[ids], reference :146-148)
            mode = self.forward_mode.value if self.fused else QuantizeForwardMode.STE.value
            emb, _res, ids, loss, _ll = ops.rq_apply(x, codebook.unsqueeze(0), mode, self.training,
                                                     self.commitment_weight, algo=self.algo,
                                                     want_residuals=False, want_level_loss=False)
            emb_out = emb[0]
            if self.training and self.forward_mode == QuantizeForwardMode.ROTATION_TRICK and x.shape[0] == 1:
                emb_out = emb_out.squeeze()  # the reference's transform ends in .squeeze() (quantize.py:45)
            return QuantizeOutput(embeddings=emb_out, ids=ids[:, 0], loss=loss)
        if (self.forward_mode == QuantizeForwardMode.GUMBEL_SOFTMAX and self.distance_mode == QuantizeDistance.L2
                and ops.gumbel_supported(self.embed_dim, self.n_embed)):
            # training, soft assignment (reference :125-130): one fused kernel, noise from Philox (seeded by torch's generator)
            emb_out, ids, loss = ops.gumbel_apply(x, codebook, float(temperature), self.commitment_weight)
            return QuantizeOutput(embeddings=emb_out, ids=ids, loss=loss)
        return self._forward_dense(x, codebook, temperature)

    def _forward_dense(self, x: Tensor, codebook: Tensor, temperature: float) -> QuantizeOutput:
        """COSINE distance, or a Gumbel shape without a fused instantiation: dense [N, K] path on PyTorch GPU ops
        (reference :108-148)."""
        if self.distance_mode == QuantizeDistance.L2:
            dist = (x ** 2).sum(dim=1, keepdim=True) + (codebook.T ** 2).sum(dim=0, keepdim=True) - 2 * x @ codebook.T
        else:
            dist = -(x / x.norm(dim=1, keepdim=True) @ codebook.T / codebook.T.norm(dim=0, keepdim=True))
        ids = dist.detach().argmin(dim=1)
        if not self.training:
            emb_out = self.get_item_embeddings(ids)
            return QuantizeOutput(emb_out, ids, self.quantize_loss(query=x, value=emb_out))
        if self.forward_mode == QuantizeForwardMode.GUMBEL_SOFTMAX:
            emb = gumbel_softmax_sample(-dist, temperature=temperature, device=self.device) @ codebook
            emb_out = emb
        else:
            emb = self.get_item_embeddings(ids)
            if self.forward_mode == QuantizeForwardMode.STE:
                emb_out = x + (emb - x).detach()
            else:
                emb_out = efficient_rotation_trick_transform(x / (x.norm(dim=-1, keepdim=True) + 1e-8),
                                                             emb / (emb.norm(dim=-1, keepdim=True) + 1e-8), x)
        return QuantizeOutput(emb_out, ids, self.quantize_loss(query=x, value=emb))
